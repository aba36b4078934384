"""profiles/r2_traffic.json from an `ncu --page raw --csv` export of tools/profile_passes.py: DRAM bytes (read + write) per launch
of the profiled kernels, tied to the kernel sources by bench.kernel_source_hash() so that bench.py refuses a stale capture.
   python tools/make_traffic_json.py gpurun_out/r2_full_raw.csv "1:2 3:2 1:1 3:0 8:0"  """
import csv, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def mb(r, k):
    v = float(r[ix[k]].replace(",", "")); u = units[ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
launches = []
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("wun::", "")
    launches.append({"kernel": name, "grid": r[ix["Grid Size"]], "us": float(r[ix["gpu__time_duration.sum"]]),
                     "dram_bytes": mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum"),
                     "tensor_pipe_active_pct": float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
# family totals over the profiled launch groups (tensor-core kernels + the split passes that feed them)
fam = {"wgrad": [l for l in launches if l["kernel"].startswith(("wgrad_umma", "split_views"))],
       "conv_dgrad": [l for l in launches if "true>" in l["kernel"] or ", 1>" in l["kernel"]],
       "conv_fwd": [l for l in launches if ("false>" in l["kernel"] or ", 0>" in l["kernel"] or "fold" in l["kernel"])]}
out = {"source_hash": bench.kernel_source_hash(), "passes": sys.argv[2] if len(sys.argv) > 2 else None,
       "note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (tools/profile_passes.py, M4 batch 16); "
               "per-family value = mean over the profiled tensor-core launches of that family",
       "launches": launches,
       "dram_bytes_per_launch": {k: (sum(l["dram_bytes"] for l in v) / len(v) if v else None) for k, v in fam.items()}}
json.dump(out, open(os.path.join(REPO, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out["dram_bytes_per_launch"]), out["source_hash"])
