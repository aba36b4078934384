"""Double-buffered host -> device input feed for the training step.

The step reads its inputs (mix, targets) from FIXED device buffers (they are baked into the CUDA graph of the step).  A
DevicePrefetcher copies the next batch from pinned host memory into one of two staging buffers on its own copy stream
while the current step computes, then hands it over with a device-to-device copy (tens of microseconds) on the compute
stream:

    pf = DevicePrefetcher([mix_dev, targets_dev])
    pf.issue([mix_pinned, targets_pinned])            # first batch
    for batch in batches:
        pf.issue(next_pinned_batch)                    # H2D of step i+1, overlaps the compute of step i
        pf.consume()                                   # step i's inputs land in mix_dev / targets_dev (compute stream)
        run_step()                                     # graph replay / loss_and_gradients + adam_step

Ordering is by CUDA events only (no host synchronisation): a staging slot is overwritten only after the D2D copy that read
it has been enqueued AND completed (`free` event), and it is consumed only after its H2D has completed (`ready` event).
This is the input half of SURVEY 8f row N4 (the reference feeds `sess.run` from a multi-threaded queue, Training.py:38-41).
"""
import torch


class DevicePrefetcher(object):
    def __init__(self, device_inputs, slots=2):
        self.inputs = list(device_inputs)
        self.copy_stream = torch.cuda.Stream(device=self.inputs[0].device)
        self.stage = [[torch.empty_like(t) for t in self.inputs] for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [torch.cuda.Event() for _ in range(slots)]
        self.n_issued = 0
        self.n_consumed = 0

    def issue(self, pinned_batch):
        """Enqueue the H2D copy of one batch (list of pinned host tensors, same shapes as the device inputs)."""
        if self.n_issued - self.n_consumed >= len(self.stage):
            raise RuntimeError("DevicePrefetcher: all staging slots are in flight (consume() first)")
        slot = self.n_issued % len(self.stage)
        self.n_issued += 1
        self.copy_stream.wait_event(self.free[slot])          # no-op until the slot has been consumed once
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.stage[slot], pinned_batch):
                dst.copy_(src, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def consume(self):
        """On the CURRENT (compute) stream: wait for the oldest issued batch and move it into the step's input buffers."""
        if self.n_consumed >= self.n_issued:
            raise RuntimeError("DevicePrefetcher: nothing issued")
        slot = self.n_consumed % len(self.stage)
        self.n_consumed += 1
        cur = torch.cuda.current_stream(self.inputs[0].device)
        cur.wait_event(self.ready[slot])
        for dst, src in zip(self.inputs, self.stage[slot]):
            dst.copy_(src, non_blocking=True)
        self.free[slot].record(cur)
