// umma_rate_bench.cu - tensor-pipe cost of one "product" K step (the 3 bf16 MMAs hi*hi + lo*hi + hi*lo of the
// fp32-accurate scheme, M = 128, K = 16, SWIZZLE_NONE K-major operands in shared memory) when the operands CHANGE from
// MMA to MMA as they do in plane_conv_umma_* (A: hi/lo sub-slabs x tap shifts x slab stages, B: a ring of weight blocks),
// next to the same-operand loop tools/umma_layout_bench measures, and the fused-N form (A_hi x [B_hi|B_lo], A_lo x B_hi).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate_bench tools/umma_rate_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t par) {
    long long t0 = clock64();
    while (!mbar_try(bar, par)) if (clock64() - t0 > 2000000000LL) asm volatile("trap;");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// mode 0: same operands every MMA; 1: rotating operands (3 MMAs per product); 2: rotating, fused-N (2 MMAs per product);
// 3: rotating, hi*hi only (1 MMA per K step: what a single-pass bf16 kernel would issue)
template <int STYLE>
__global__ void __launch_bounds__(384, 1) rate_kernel(int mode, int N, int MT, int iters, long long* out, const int* __restrict__ taps, int bg, volatile int* stop_flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_holder)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_holder;
    if (warp >= 4) {
        // background shared-memory traffic (what converter / epilogue warps generate next to the MMA issuer):
        // bg 1: 16-byte stores, bg 2: 16-byte loads, bg 3: both; into the last 8 KB of the buffer (not an MMA operand)
        uint4* scratch = reinterpret_cast<uint4*>(smem + 192 * 1024) + (threadIdx.x - 128);
        uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
        if (bg) {
            while (!done) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (bg & 1) scratch[(r & 1) * 256] = v;
                    if (bg & 2) { uint4 t = scratch[(r & 1) * 256]; v.x ^= t.y; }
                }
            }
        }
        if (v.x == 0xdeadbeef) out[1] = v.x;
    }
    if (STYLE == 0) {
    if (warp == 0 && elect_one()) {        // elect.sync, not lane == 0: otherwise every MMA is wrapped in an ELECT retry loop
        // slab stage: [hi a0 | hi a1 | lo a0 | lo a1][rows_alloc][16 B]; 3 stages; weight ring behind them: blocks of 64*N bytes
        const uint32_t rows_alloc = MT * 128 + 8, atom = 16u * rows_alloc, slab = 64u * rows_alloc;
        const uint32_t ring0 = smem_u32(smem) + 3 * slab, blk = 64u * N;
        const int nblk = (int)((190u * 1024u - 3 * slab) / blk);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 2) << 17) | ((128u >> 4) << 24);
        long long t0 = clock64();
        int b = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t sa = smem_u32(smem) + (mode ? (uint32_t)(i % 3) * slab : 0u);
            const uint64_t a_hi0 = make_desc(sa, atom, 128), a_lo0 = make_desc(sa + 2 * atom, atom, 128);
            for (int tap = 0; tap < 8; ++tap) {
                const uint32_t sb = ring0 + (mode ? (uint32_t)b * blk : 0u);
                if (++b >= nblk) b = 0;
                const uint64_t b_hi = make_desc(sb, 32u * N, 128), b_lo = make_desc(sb + 16u * N, 32u * N, 128);
                for (int mt = 0; mt < MT; ++mt) {
                    const uint64_t sh = mode ? (uint64_t)(taps[tap] + 128 * mt) : 0ull;
                    const uint32_t td = tm + (uint32_t)(mt * ((mode == 2) ? 2 * N : N));
                    if (mode == 2) { mma(td, a_hi0 + sh, b_hi, idesc2, 1); mma(td, a_lo0 + sh, b_hi, idesc, 1); }
                    else if (mode == 3) { mma(td, a_hi0 + sh, b_hi, idesc, 1); }
                    else { mma(td, a_lo0 + sh, b_hi, idesc, 1); mma(td, a_hi0 + sh, b_lo, idesc, 1); mma(td, a_hi0 + sh, b_hi, idesc, 1); }
                }
            }
        }
        commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        done = 1;
    }
    } else if (warp == 0) {
        // STYLE 1: every lane of the issuing warp runs the loop (warp-uniform values -> uniform datapath), only the MMAs of
        // one tap sit in an elected region; MT is a compile-time-like switch (two straight-line bodies), descriptors advance
        // by adds, the tap shifts come from a small table read ahead.
        const uint32_t rows_alloc = MT * 128 + 8, atom = 16u * rows_alloc, slab = 64u * rows_alloc;
        const uint32_t ring0 = smem_u32(smem) + 3 * slab, blk = 64u * N;
        const int nblk = (int)((190u * 1024u - 3 * slab) / blk);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 2) << 17) | ((128u >> 4) << 24);
        const uint64_t lo_off = (uint64_t)((2 * atom) >> 4), blo_off = (uint64_t)((16u * N) >> 4), blk_off = (uint64_t)(blk >> 4);
        const uint32_t acc2 = (mode == 2) ? 2 * N : N;
        long long t0 = clock64();
        int b = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t sa = smem_u32(smem) + (mode ? (uint32_t)(i % 3) * slab : 0u);
            const uint64_t a_hi0 = make_desc(sa, atom, 128);
            if (b + 8 > nblk) b = 0;                                  // ring wrap once per job, not per tap
            uint64_t b_hi = make_desc(ring0 + (mode ? (uint32_t)b * blk : 0u), 32u * N, 128);
            b += 8;
#pragma unroll
            for (int tap = 0; tap < 8; ++tap) {
                const uint64_t a_hi = a_hi0 + (uint64_t)(mode ? taps[tap] : 0), a_lo = a_hi + lo_off, b_lo = b_hi + blo_off;
                if (elect_one()) {
                    if (mode == 2) { mma(tm, a_hi, b_hi, idesc2, 1); mma(tm, a_lo, b_hi, idesc, 1); }
                    else if (mode == 3) { mma(tm, a_hi, b_hi, idesc, 1); }
                    else { mma(tm, a_lo, b_hi, idesc, 1); mma(tm, a_hi, b_lo, idesc, 1); mma(tm, a_hi, b_hi, idesc, 1); }
                    if (MT == 2) {
                        if (mode == 2) { mma(tm + acc2, a_hi + 128, b_hi, idesc2, 1); mma(tm + acc2, a_lo + 128, b_hi, idesc, 1); }
                        else if (mode == 3) { mma(tm + acc2, a_hi + 128, b_hi, idesc, 1); }
                        else { mma(tm + acc2, a_lo + 128, b_hi, idesc, 1); mma(tm + acc2, a_hi + 128, b_lo, idesc, 1); mma(tm + acc2, a_hi + 128, b_hi, idesc, 1); }
                    }
                }
                __syncwarp();
                if (mode) b_hi += blk_off;
            }
        }
        if (elect_one()) commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
        done = 1;
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* dout; CK(cudaMalloc(&dout, 64));
    CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int htaps[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    int* dtaps; CK(cudaMalloc(&dtaps, sizeof(htaps))); CK(cudaMemcpy(dtaps, htaps, sizeof(htaps), cudaMemcpyHostToDevice));
    const char* names[4] = {"same operands, 3 MMA", "rotating operands, 3 MMA", "rotating, fused-N 2 MMA", "rotating, 1 MMA (hi*hi)"};
    printf("cycles per PRODUCT K step (M=128 rows x N x K=16; math floor at 4096 MAC/clk: 3 MMAs = 1.5*N, fused = 1.5*N)\n");
    for (int N : {32, 48, 96, 144})
        for (int MT : {1, 2}) {
            for (int mode : {1, 2}) {
                if (mode == 2 && 2 * N > 256) continue;
                const int iters = 60;
                double r[4];
                for (int bg = 0; bg < 4; ++bg) {
                    rate_kernel<0><<<148, 384, 200 * 1024>>>(mode, N, MT, iters, dout, dtaps, bg, nullptr);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
                    long long cyc; CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
                    r[bg] = (double)cyc / (iters * 8 * MT);
                }
                printf("N=%3d MT=%d %-26s : alone %6.1f | +8 warps storing %6.1f | loading %6.1f | both %6.1f   cycles / K step / row tile\n",
                       N, MT, names[mode], r[0], r[1], r[2], r[3]);
            }
        }
    return 0;
}
