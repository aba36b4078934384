// presplit_probe.cu - ROUND-2 EXPERIMENT (written at the end of round 1 with no GPU time left: compiles, NOT yet run).
//
// Question: how fast is the tcgen05 plane convolution when the A operand is NOT converted fp32 -> (hi, lo) bf16 by
// converter warps in every consumer CTA, but already lies in HBM in slab layout and is pulled into shared memory by
// cp.async.bulk (DESIGN.md section 8.1 "split once, not per consumer")?
//
//   split storage of one plane:  [batch][16-channel chunk][hi a0 | hi a1 | lo a0 | lo a1][row][8 x bf16 = 16 B]
//   slab of (tile, chunk)      =  4 contiguous row ranges of that array -> 4 bulk copies of rows_alloc*16 B, issued by one
//                                 thread, landing in exactly the K-major SWIZZLE_NONE layout the MMA descriptors expect.
//
// The probe runs the SAME down-block shaped problem through (a) the engine's kernel (umma_run_conv path) and (b) the
// bulk-fed variant below (same weight packs, same descriptors, same double-buffered-TMEM epilogue), checks both against
// the CPU reference and prints both times.  Usage (under gpurun):   tools/presplit_probe [all|small|down1..down4|wsmall|wdown1|wdown2|wdown3|wdown5]
// (the w* cases do the same for the wgrad kernel: presplit_wgrad_kernel vs wgrad_umma_kernel)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../wave-u-net_b200/csrc/kernels_umma.cu"

using namespace wun;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// fp32 plane -> split storage (in the redesign this is the producer's epilogue)
// ------------------------------------------------------------------------------------------------
struct SplitPlanes {
    const uint8_t* base[kMaxPlanes];     // per plane
    long long bstride, cstride, pstride; // bytes: batch, 16-channel chunk, sub-plane (= Rpad * 16)
    int nchunk;
};

__global__ void split_plane_kernel(PlaneView P, int B, uint8_t* out, int nchunk, int Rpad) {
    const long long total = (long long)B * nchunk * Rpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % Rpad);
        const int c = (int)((i / Rpad) % nchunk);
        const int b = (int)(i / ((long long)Rpad * nchunk));
        float x[16];
        load_row16(P, b, r, c * 16, x);              // zero outside the valid rows / channels
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * k]), h1 = __float2bfloat16_rn(x[2 * k + 1]);
            hi[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[k] = pack_bf16x2(x[2 * k] - __bfloat162float(h0), x[2 * k + 1] - __bfloat162float(h1));
        }
        uint8_t* o = out + (((long long)b * nchunk + c) * 4) * (long long)Rpad * 16 + (long long)r * 16;
        const long long ps = (long long)Rpad * 16;
        *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(o + ps) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(o + 2 * ps) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(o + 3 * ps) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
}

// ------------------------------------------------------------------------------------------------
// bulk-fed persistent conv: warp 0 slab loader, warp 1 weight loader, warp 2 TMEM alloc + MMA issue, warp 3 idle,
// warps 4-7 epilogue (TMEM lane quarter = warp & 3).  NS slab stages.  Everything after the slab fill is the engine's
// persistent kernel (plane_conv_umma_persistent): same descriptors, weight ring, double-buffered accumulators, epilogue.
// ------------------------------------------------------------------------------------------------
constexpr int kPsSlabStages = 4;
constexpr int kPsThreads = 256;

// per-CTA cycle counters of the MMA-issuing thread (always compiled in: five clock64 reads per K chunk are noise):
// [0] CTAs, [1] whole tile loop, [2] waiting for a slab, [3] waiting for weights, [4] waiting for a free accumulator
__device__ unsigned long long g_ps_timing[8];

// split OUTPUT storage (what the next layer's slab loader wants): per class, row m of the class goes to parity plane m & 1
// at index m >> 1:  [class][parity][batch][16-channel chunk][hi a0 | hi a1 | lo a0 | lo a1][row][16 B]
struct SplitOut {
    uint8_t* base[kMaxClasses][2];
    long long bstride, cstride, pstride;
};

// SPLIT_OUT = false: fp32 row-major output through the engine's shared-memory staged epilogue.
// SPLIT_OUT = true : thread = accumulator row writes hi/lo bf16 atoms straight from registers (no staging tile).
template <bool SPLIT_OUT>
__global__ void __launch_bounds__(kPsThreads, 1) presplit_conv_persistent(const __grid_constant__ UmmaLaunch L, int total_tiles,
                                                                          const __grid_constant__ SplitPlanes XS,
                                                                          const __grid_constant__ SplitOut YS) {
    constexpr int NS = kPsSlabStages;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NPAD = L.NPAD;
    const uint32_t slab_bytes = 64u * L.rows_alloc;
    const uint32_t bblk_bytes = 64u * NPAD;
    const int TB = L.TB, nbs = L.nbs;
    const uint32_t bstage_bytes = bblk_bytes * TB;
    const int CW = (NPAD < 128) ? NPAD : 128;
    const int SW = CW + 4;
    uint8_t* slab0 = smem;
    uint8_t* bring0 = smem + NS * slab_bytes;
    float* stage = reinterpret_cast<float*>(bring0 + nbs * bstage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + (size_t)128 * SW * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int SLAB_FULL = 0, SLAB_EMPTY = NS, B_FULL = 2 * NS, B_EMPTY = 2 * NS + kBStagesMax, ACC_FULL = 2 * NS + 2 * kBStagesMax,
              ACC_EMPTY = ACC_FULL + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + ACC_EMPTY + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_holder + 4);
    for (int i = tid; i < NPAD * L.nsplit; i += blockDim.x) bias_s[i] = (L.bias && i < L.N) ? __ldg(L.bias + i) : 0.f;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(BAR(SLAB_FULL + i), 1); mbar_init(BAR(SLAB_EMPTY + i), 1); }
        for (int i = 0; i < kBStagesMax; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(ACC_FULL + i), 1); mbar_init(BAR(ACC_EMPTY + i), 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_holder), L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t acc_cols = (uint32_t)(L.MT * NPAD);
    const uint32_t atom_stride = 16u * L.rows_alloc;

    if (warp == 0) {
        // ===================== slab loader: 4 bulk copies per (tile, group, chunk) =====================
        if (elect_one()) {
            int jg = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(L, t);
                const UmmaClass& K = L.cls[tc.cls];
                for (int g = 0; g < K.ngroups; ++g) {
                    const UmmaGroup& G = K.groups[g];
                    const int nchunk = (L.planes[G.plane].C + 15) >> 4;
                    for (int c = 0; c < nchunk; ++c, ++jg) {
                        const int st = jg % NS;
                        mbar_wait(BAR(SLAB_EMPTY + st), ((jg / NS) & 1) ^ 1);
                        mbar_arrive_expect_tx(BAR(SLAB_FULL + st), 4u * atom_stride);
                        const uint8_t* src = XS.base[G.plane] + (long long)tc.b * XS.bstride + (long long)c * XS.cstride +
                                             (long long)(tc.m_base + G.dmin) * 16;
                        const uint32_t dst = smem_u32(slab0 + st * slab_bytes);
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            bulk_g2s(dst + p * atom_stride, src + (long long)p * XS.pstride, atom_stride, BAR(SLAB_FULL + st));
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== weight loader (as in the engine) =====================
        if (elect_one()) {
            int bg = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(L, t);
                const UmmaClass& K = L.cls[tc.cls];
                const uint8_t* src = K.wpack[tc.split];
                size_t blk = 0;
                for (int g = 0; g < K.ngroups; ++g) {
                    const UmmaGroup& G = K.groups[g];
                    const int nchunk = (L.planes[G.plane].C + 15) >> 4;
                    const int nterm = G.term_end - G.term_begin;
                    for (int c = 0; c < nchunk; ++c)
                        for (int t0 = 0; t0 < nterm; t0 += TB, ++bg) {
                            const int nt = min(TB, nterm - t0);
                            const int bs = bg % nbs;
                            mbar_wait(BAR(B_EMPTY + bs), ((bg / nbs) & 1) ^ 1);
                            mbar_arrive_expect_tx(BAR(B_FULL + bs), bblk_bytes * nt);
                            bulk_g2s(smem_u32(bring0 + bs * bstage_bytes), src + blk * bblk_bytes, bblk_bytes * nt, BAR(B_FULL + bs));
                            blk += nt;
                        }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== MMA issuer (as in the engine) =====================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t b_lbo = 32u * NPAD;
            int jg = 0, bg = 0, k = 0;
            long long t_slab = 0, t_b = 0, t_acc = 0;
            const long long t_loop0 = clock64();
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
                const TileCoord tc = decode_tile(L, t);
                const UmmaClass& K = L.cls[tc.cls];
                const int buf = k & 1;
                { const long long w0 = clock64(); mbar_wait(BAR(ACC_EMPTY + buf), ((k >> 1) & 1) ^ 1); t_acc += clock64() - w0; }
                tc_fence_after();
                uint32_t first = 0;
                for (int g = 0; g < K.ngroups; ++g) {
                    const UmmaGroup& G = K.groups[g];
                    const int nchunk = (L.planes[G.plane].C + 15) >> 4;
                    for (int c = 0; c < nchunk; ++c, ++jg) {
                        const int st = jg % NS;
                        { const long long w0 = clock64(); mbar_wait(BAR(SLAB_FULL + st), (jg / NS) & 1); t_slab += clock64() - w0; }
                        tc_fence_after();
                        const uint32_t sa = smem_u32(slab0 + st * slab_bytes);
                        const uint64_t a_hi0 = umma_desc(sa, atom_stride, 128), a_lo0 = umma_desc(sa + 2 * atom_stride, atom_stride, 128);
                        for (int t0 = G.term_begin; t0 < G.term_end; t0 += TB, ++bg) {
                            const int bs = bg % nbs;
                            { const long long w0 = clock64(); mbar_wait(BAR(B_FULL + bs), (bg / nbs) & 1); t_b += clock64() - w0; }
                            tc_fence_after();
                            const uint32_t sb = smem_u32(bring0 + bs * bstage_bytes);
                            const uint64_t b_hi0 = umma_desc(sb, b_lbo, 128), b_lo0 = umma_desc(sb + 16u * NPAD, b_lbo, 128);
                            const int nt = min(TB, G.term_end - t0);
                            for (int tt = 0; tt < nt; ++tt) {
                                const uint64_t boff = (uint64_t)((bblk_bytes >> 4) * tt);
                                const uint64_t b_hi = b_hi0 + boff, b_lo = b_lo0 + boff;
                                const uint64_t aoff = (uint64_t)(uint32_t)(L.d[t0 + tt] - G.dmin);
                                for (int mt = 0; mt < L.MT; ++mt) {
                                    const uint64_t a_hi = a_hi0 + aoff + (uint64_t)(128u * mt), a_lo = a_lo0 + aoff + (uint64_t)(128u * mt);
                                    const uint32_t td = tmem_base + buf * acc_cols + (uint32_t)(mt * NPAD);
                                    umma_bf16(td, a_lo, b_hi, idesc, first);
                                    umma_bf16(td, a_hi, b_lo, idesc, 1u);
                                    umma_bf16(td, a_hi, b_hi, idesc, 1u);
                                }
                                first = 1u;
                            }
                            umma_commit(BAR(B_EMPTY + bs));
                        }
                        umma_commit(BAR(SLAB_EMPTY + st));
                    }
                }
                umma_commit(BAR(ACC_FULL + buf));
            }
            atomicAdd(&g_ps_timing[0], 1ull);
            atomicAdd(&g_ps_timing[1], (unsigned long long)(clock64() - t_loop0));
            atomicAdd(&g_ps_timing[2], (unsigned long long)t_slab);
            atomicAdd(&g_ps_timing[3], (unsigned long long)t_b);
            atomicAdd(&g_ps_timing[4], (unsigned long long)t_acc);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue warps 4-7 (as in the engine; fp32 row-major output, forward epilogue only) ==========
        const int q4 = warp & 3;
        int k = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
            const TileCoord tc = decode_tile(L, t);
            const UmmaClass& K = L.cls[tc.cls];
            const int buf = k & 1;
            const int n0 = tc.split * NPAD;
            mbar_wait(BAR(ACC_FULL + buf), (k >> 1) & 1);
            tc_fence_after();
            if (SPLIT_OUT) {
                for (int mt = 0; mt < L.MT; ++mt) {
                    const int m = tc.m_base + mt * 128 + q4 * 32 + lane;
                    const bool row_ok = m < K.out.m_hi;
                    for (int cb = 0; cb < NPAD; cb += 16) {
                        if (n0 + cb >= ((L.N + 15) & ~15)) break;
                        __syncwarp();
                        float v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * acc_cols + (uint32_t)(mt * NPAD + cb), v);
                        if (!row_ok) continue;
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float y0 = v[2 * j] + bias_s[n0 + cb + 2 * j], y1 = v[2 * j + 1] + bias_s[n0 + cb + 2 * j + 1];
                            y0 = fmaxf(0.2f * y0, y0); y1 = fmaxf(0.2f * y1, y1);
                            const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
                            hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo[j] = pack_bf16x2(y0 - __bfloat162float(h0), y1 - __bfloat162float(h1));
                        }
                        uint8_t* o = YS.base[tc.cls][m & 1] + (long long)tc.b * YS.bstride + (long long)((n0 + cb) >> 4) * YS.cstride +
                                     (long long)(m >> 1) * 16;
                        *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(o + YS.pstride) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                        *reinterpret_cast<uint4*>(o + 2 * YS.pstride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<uint4*>(o + 3 * YS.pstride) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
                tc_fence_before();
                mbar_arrive(BAR(ACC_EMPTY + buf));
                continue;
            }
            const int c0_last = min((NPAD - 1) / CW, (L.N - n0 - 1) / CW) * CW;
            for (int mt = 0; mt < L.MT; ++mt) {
                for (int c0 = 0; c0 < NPAD; c0 += CW) {
                    const bool last_block = (mt == L.MT - 1) && (c0 == c0_last);
                    if (n0 + c0 < L.N) {
                        const int cw = min(CW, NPAD - c0);
                        for (int cb = 0; cb < cw; cb += 16) {
                            __syncwarp();
                            float v[16];
                            tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * acc_cols + (uint32_t)(mt * NPAD + c0 + cb), v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float y = v[j] + bias_s[n0 + c0 + cb + j];
                                v[j] = fmaxf(0.2f * y, y);
                            }
                            float4* dst = reinterpret_cast<float4*>(stage + (size_t)(q4 * 32 + lane) * SW + cb);
#pragma unroll
                            for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    }
                    if (last_block) {
                        tc_fence_before();
                        mbar_arrive(BAR(ACC_EMPTY + buf));
                    }
                    if (n0 + c0 < L.N) {
                        const int cw = min(CW, NPAD - c0);
                        __syncwarp();
                        const int ncols = min(cw, L.N - (n0 + c0));
                        const long long tile_off = (long long)tc.b * K.out.bstride + n0 + c0;
                        if (ncols % 4 == 0 && K.out.rstride % 4 == 0) {
                            const int Q = ncols >> 2;
                            for (int it = lane; it < 32 * Q; it += 32) {
                                const int rl = it / Q, q = it - rl * Q;
                                const int r = q4 * 32 + rl;
                                const int m = tc.m_base + mt * 128 + r;
                                if (m >= K.out.m_hi) continue;
                                const long long roff = tile_off + (long long)m * K.out.rstride;
                                *(reinterpret_cast<float4*>(K.out.base + roff) + q) = *reinterpret_cast<const float4*>(stage + (size_t)r * SW + 4 * q);
                            }
                        } else {
                            for (int it = lane; it < 32 * ncols; it += 32) {
                                const int rl = it / ncols, j = it - rl * ncols;
                                const int r = q4 * 32 + rl;
                                const int m = tc.m_base + mt * 128 + r;
                                if (m >= K.out.m_hi) continue;
                                K.out.base[tile_off + (long long)m * K.out.rstride + j] = stage[(size_t)r * SW + j];
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, L.tmem_cols);
    }
}

static size_t presplit_smem_bytes(const UmmaLaunch& L) {
    const size_t CW = (L.NPAD < 128) ? L.NPAD : 128;
    return (size_t)kPsSlabStages * 64u * L.rows_alloc + (size_t)L.nbs * L.TB * 64u * L.NPAD + 128u * (CW + 4) * 4 +
           (2 * kPsSlabStages + 2 * kBStagesMax + 4) * 8 + 32 + 4 * (size_t)L.NPAD * L.nsplit;
}

struct Problem {
    int B, T, Cin, Cout, fs;
    int cs, U;      // skip window [cs, cs+U) in full-rate output coordinates
};

static void run_case(const char* name, Problem p, bool check, int timing_iters) {
    const int To = p.T - p.fs + 1, Td = (To + 1) / 2;
    const int mo_lo = p.cs / 2, mo_hi = (p.cs + p.U) / 2, n_odd = mo_hi - mo_lo;
    std::vector<float> x((size_t)p.B * p.T * p.Cin), w((size_t)p.fs * p.Cin * p.Cout), bias(p.Cout);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : x) v = rnd();
    for (auto& v : w) v = rnd() * 0.2f;
    for (auto& v : bias) v = rnd() * 0.1f;
    float *dx, *dw, *db, *ddec, *dodd;
    const size_t dec_bytes = (size_t)p.B * Td * p.Cout * 4, odd_bytes = (size_t)p.B * (n_odd + 1) * p.Cout * 4;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4)); CK(cudaMalloc(&db, bias.size() * 4));
    CK(cudaMalloc(&ddec, dec_bytes)); CK(cudaMalloc(&dodd, odd_bytes));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));

    // the down-block launch exactly as the engine builds it: two parity planes of the input, classes dec / odd
    ConvLaunch CL;
    memset(&CL, 0, sizeof(CL));
    CL.nplanes = 2;
    for (int par = 0; par < 2; ++par) {
        PlaneView& P = CL.planes[par];
        P.base = dx + par * p.Cin; P.bstride = (long long)p.T * p.Cin; P.rstride = 2 * p.Cin;
        P.r_lo = 0; P.r_hi = (par == 0) ? (p.T + 1) / 2 : p.T / 2; P.C = p.Cin; P.kind = PLANE_DIRECT;
    }
    CL.ncls = 2; CL.N = p.Cout; CL.w_sk = p.Cout; CL.w_sn = 1; CL.W = dw; CL.bias = db;
    CL.epilogue = EPI_BIAS_LRELU; CL.batch = p.B;
    int nt = 0;
    for (int q = 0; q < 2; ++q) {
        OutView& O = CL.cls[q];
        O.base = (q == 0) ? ddec : dodd - (long long)mo_lo * p.Cout;
        O.bstride = (q == 0) ? (long long)Td * p.Cout : (long long)n_odd * p.Cout;
        O.rstride = p.Cout;
        O.m_lo = (q == 0) ? 0 : mo_lo; O.m_hi = (q == 0) ? Td : mo_hi;
        O.term_begin = nt;
        for (int par = 0; par < 2; ++par)
            for (int j = 0; j < p.fs; ++j) {
                int e = q + j;
                if ((e & 1) != par) continue;
                CL.terms[nt++] = {par, e >> 1, j * p.Cin * p.Cout};
            }
        O.term_end = nt;
        CL.max_rows = max(CL.max_rows, O.m_hi - O.m_lo);
    }
    UmmaChoice ch;
    if (!umma_plan_from_conv(CL, &ch)) { printf("[%s] not eligible\n", name); exit(4); }
    uint8_t* arena;
    CK(cudaMalloc(&arena, ch.pack_bytes));
    UmmaLaunch L;
    UmmaPackLaunch PL;
    CK(umma_build(CL, ch, arena, &L, &PL));
    CK(launch_umma_pack(PL, 0));
    CK(cudaDeviceSynchronize());
    printf("[%s] B=%d T=%d Cin=%d Cout=%d fs=%d | engine kernel: persistent=%d MT=%d NPAD=%d nsplit=%d TB=%d nbs=%d\n", name, p.B, p.T,
           p.Cin, p.Cout, p.fs, L.persistent, L.MT, L.NPAD, L.nsplit, L.TB, L.nbs);

    // ---- bulk-fed variant: same packs; MT = 2 tile pairs, double-buffered accumulators ----
    UmmaLaunch LB = L;
    LB.persistent = 1;
    if (2 * LB.MT * LB.NPAD > 512) { LB.MT = 1; LB.rows_alloc = 128 + (L.rows_alloc - L.MT * 128); }
    { int tm = 32; while (tm < 2 * LB.MT * LB.NPAD) tm *= 2; LB.tmem_cols = tm; }
    if (LB.tmem_cols > 512) { printf("[%s] NPAD too wide for double-buffered TMEM - skipping the bulk-fed variant\n", name); return; }
    {   // weight ring: what fits next to 4 slab stages and the staging tile in ~218 KB
        const long long cw = (LB.NPAD < 128) ? LB.NPAD : 128;
        long long left = 218 * 1024 - (long long)kPsSlabStages * 64 * LB.rows_alloc - 128 * (cw + 4) * 4 - 1024 - 4LL * LB.NPAD * LB.nsplit;
        if (left > 98304) left = 98304;
        LB.nbs = (int)(left / ((long long)LB.TB * 64 * LB.NPAD));
        if (LB.nbs > kBStagesMax) LB.nbs = kBStagesMax;
        if (LB.nbs < 2) { printf("[%s] weight ring does not fit\n", name); return; }
    }
    const int nchunk = (p.Cin + 15) / 16;
    const int Rpad = (p.T + 1) / 2 + LB.rows_alloc + 64;          // zero rows behind every plane: bulk copies never leave the array
    SplitPlanes XS;
    memset(&XS, 0, sizeof(XS));
    XS.nchunk = nchunk; XS.pstride = (long long)Rpad * 16; XS.cstride = 4 * XS.pstride; XS.bstride = nchunk * XS.cstride;
    uint8_t* dxs[2];
    for (int par = 0; par < 2; ++par) {
        const size_t bytes = (size_t)p.B * XS.bstride;
        CK(cudaMalloc(&dxs[par], bytes));
        split_plane_kernel<<<148 * 8, 256>>>(CL.planes[par], p.B, dxs[par], nchunk, Rpad);
        CK(cudaGetLastError());
        XS.base[par] = dxs[par];
    }
    CK(cudaDeviceSynchronize());
    int total = 0;
    for (int q = 0; q < LB.ncls; ++q) total += ((LB.cls[q].out.m_hi - LB.cls[q].out.m_lo + LB.MT * 128 - 1) / (LB.MT * 128)) * LB.batch;
    total *= LB.nsplit;
    const size_t smem_b = presplit_smem_bytes(LB);
    CK(cudaFuncSetAttribute(presplit_conv_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    printf("[%s] bulk-fed variant: MT=%d rows_alloc=%d tmem=%d nbs=%d slab stages=%d smem=%zu tiles=%d\n", name, LB.MT, LB.rows_alloc,
           LB.tmem_cols, LB.nbs, kPsSlabStages, smem_b, total);
    const int grid = total < 148 ? total : 148;
    // split output arrays (class q, parity): rows of class q = [m_lo, m_hi) -> index m >> 1
    SplitOut YS;
    memset(&YS, 0, sizeof(YS));
    const int nchunk_o = (p.Cout + 15) / 16;
    const int Rpad_o = (max(Td, mo_hi) + 1) / 2 + 8;
    YS.pstride = (long long)Rpad_o * 16; YS.cstride = 4 * YS.pstride; YS.bstride = nchunk_o * YS.cstride;
    const size_t ys_bytes = (size_t)p.B * YS.bstride;
    for (int q = 0; q < 2; ++q)
        for (int par = 0; par < 2; ++par) { CK(cudaMalloc(&YS.base[q][par], ys_bytes)); CK(cudaMemset(YS.base[q][par], 0, ys_bytes)); }
    CK(cudaFuncSetAttribute(presplit_conv_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));

    auto check_outputs = [&](const char* which, bool from_split = false) {
        std::vector<float> dec((size_t)p.B * Td * p.Cout), odd((size_t)p.B * (n_odd + 1) * p.Cout);
        if (!from_split) {
            CK(cudaMemcpy(dec.data(), ddec, dec.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(odd.data(), dodd, odd.size() * 4, cudaMemcpyDeviceToHost));
        } else {
            // rebuild fp32 rows from the hi/lo atom planes
            std::vector<uint16_t> buf(ys_bytes / 2);
            for (int q = 0; q < 2; ++q)
                for (int par = 0; par < 2; ++par) {
                    CK(cudaMemcpy(buf.data(), YS.base[q][par], ys_bytes, cudaMemcpyDeviceToHost));
                    const int m_lo = (q == 0) ? 0 : mo_lo, m_hi = (q == 0) ? Td : mo_hi;
                    for (int b = 0; b < p.B; ++b)
                        for (int m = m_lo + ((m_lo & 1) != par); m < m_hi; m += 2)
                            for (int n = 0; n < p.Cout; ++n) {
                                const size_t e0 = ((size_t)b * YS.bstride + (size_t)(n >> 4) * YS.cstride) / 2;
                                const size_t at = (size_t)((n >> 3) & 1) * (YS.pstride / 2) + (size_t)(m >> 1) * 8 + (n & 7);
                                auto f = [&](uint16_t h) { uint32_t u = (uint32_t)h << 16; float r; memcpy(&r, &u, 4); return r; };
                                const float val = f(buf[e0 + at]) + f(buf[e0 + 2 * (YS.pstride / 2) + at]);
                                if (q == 0) dec[((size_t)b * Td + m) * p.Cout + n] = val;
                                else odd[((size_t)b * n_odd + (m - mo_lo)) * p.Cout + n] = val;
                            }
                }
        }
        std::vector<float> xh(x.size()), xl(x.size()), wh(w.size()), wl(w.size());
        for (size_t i = 0; i < x.size(); ++i) { xh[i] = bf16_round(x[i]); xl[i] = bf16_round(x[i] - xh[i]); }
        for (size_t i = 0; i < w.size(); ++i) { wh[i] = bf16_round(w[i]); wl[i] = bf16_round(w[i] - wh[i]); }
        long long nbad = 0, ntested = 0;
        double worst = 0;
        const int stride = (To > 4000) ? 37 : 1;
        for (int b = 0; b < p.B; ++b)
            for (int a = 0; a < To; a += stride) {
                const bool even = (a & 1) == 0;
                if (!even && !(a >= p.cs && a < p.cs + p.U)) continue;
                for (int n = 0; n < p.Cout; ++n) {
                    double acc = bias[n];
                    for (int j = 0; j < p.fs; ++j)
                        for (int c = 0; c < p.Cin; ++c) {
                            size_t xi = ((size_t)b * p.T + a + j) * p.Cin + c, wi = ((size_t)j * p.Cin + c) * p.Cout + n;
                            acc += (double)xh[xi] * wh[wi] + (double)xl[xi] * wh[wi] + (double)xh[xi] * wl[wi];
                        }
                    double ref = acc > 0 ? acc : 0.2 * acc;
                    float got = even ? dec[((size_t)b * Td + a / 2) * p.Cout + n]
                                     : odd[((size_t)b * n_odd + ((a - 1) / 2 - mo_lo)) * p.Cout + n];
                    double err = fabs(got - ref) / (fabs(ref) + 1e-2);
                    if (!(err < 4e-4)) { if (nbad < 5) printf("   mismatch b=%d a=%d n=%d got=%g ref=%g\n", b, a, n, got, ref); ++nbad; }
                    if (err > worst || err != err) worst = err;
                    ++ntested;
                }
            }
        printf("[%s] %-9s %s  tested=%lld bad=%lld worst_rel=%.3e\n", name, which, nbad == 0 ? "PASS" : "FAIL", ntested, nbad, worst);
    };
    auto time_it = [&](const char* which, auto launch) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < timing_iters; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= timing_iters;
        const double flops = 2.0 * p.B * ((double)Td + n_odd) * p.fs * p.Cin * p.Cout;
        printf("[%s] %-9s time %.1f us  %.1f useful TFLOP/s\n", name, which, ms * 1e3, flops / (ms * 1e-3) * 1e-12);
    };

    // (a) engine kernel
    CK(cudaMemset(ddec, 0xFF, dec_bytes)); CK(cudaMemset(dodd, 0xFF, odd_bytes));
    CK(launch_plane_conv_umma(L, 0));
    { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("[%s] engine KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (check) check_outputs("engine");
    // (b) bulk-fed
    CK(cudaMemset(ddec, 0xFF, dec_bytes)); CK(cudaMemset(dodd, 0xFF, odd_bytes));
    presplit_conv_persistent<false><<<grid, kPsThreads, smem_b>>>(LB, total, XS, YS);
    { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[%s] bulk-fed KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (check) check_outputs("bulk-fed");
    // (c) bulk-fed + split-format epilogue
    presplit_conv_persistent<true><<<grid, kPsThreads, smem_b>>>(LB, total, XS, YS);
    { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[%s] bulk-fed/split-out KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (check) check_outputs("split-out", true);
    if (timing_iters > 0) {
        time_it("engine", [&]() { CK(launch_plane_conv_umma(L, 0)); });
        { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; CK(cudaMemcpyToSymbol(g_ps_timing, z, sizeof(z))); }
        time_it("bulk-fed", [&]() { presplit_conv_persistent<false><<<grid, kPsThreads, smem_b>>>(LB, total, XS, YS); });
        auto report = [&](const char* which) {
            unsigned long long tt[8];
            CK(cudaMemcpyFromSymbol(tt, g_ps_timing, sizeof(tt)));
            const double n = (double)tt[0];
            if (n > 0) printf("[%s] %-9s MMA thread per CTA: loop %.0f cycles = issue %.0f + wait slab %.0f + wait weights %.0f + wait accumulator %.0f\n",
                              name, which, tt[1] / n, (tt[1] - tt[2] - tt[3] - tt[4]) / n, tt[2] / n, tt[3] / n, tt[4] / n);
            memset(tt, 0, sizeof(tt));
            CK(cudaMemcpyToSymbol(g_ps_timing, tt, sizeof(tt)));
        };
        report("bulk-fed");
        time_it("split-out", [&]() { presplit_conv_persistent<true><<<grid, kPsThreads, smem_b>>>(LB, total, XS, YS); });
        report("split-out");
        time_it("split x2", [&]() { for (int par = 0; par < 2; ++par) split_plane_kernel<<<148 * 8, 256>>>(CL.planes[par], p.B, dxs[par], nchunk, Rpad); });
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(ddec); cudaFree(dodd); cudaFree(arena);
    for (int par = 0; par < 2; ++par) cudaFree(dxs[par]);
    for (int q = 0; q < 2; ++q) for (int par = 0; par < 2; ++par) cudaFree(YS.base[q][par]);
}


// ------------------------------------------------------------------------------------------------
// bulk-fed wgrad: the LIBRARY's split pass + wgrad_umma_bulk_kernel (kernels_umma.cu, engine switch WUN_BULK_WGRAD=1)
// next to the converter-fed wgrad_umma_kernel on the same launch description.
// ------------------------------------------------------------------------------------------------
static void run_wgrad(const char* name, Problem p, bool check, int timing_iters) {
    const int To = p.T - p.fs + 1, Td = (To + 1) / 2;
    const int mo_lo = p.cs / 2, mo_hi = (p.cs + p.U) / 2, n_odd = mo_hi - mo_lo;
    std::vector<float> x((size_t)p.B * p.T * p.Cin), gdec((size_t)p.B * Td * p.Cout), godd((size_t)p.B * (n_odd + 1) * p.Cout);
    unsigned s = 777u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : x) v = rnd();
    for (auto& v : gdec) v = rnd() * 1e-3f;
    for (auto& v : godd) v = rnd() * 1e-3f;
    float *dx, *dgd, *dgo, *ddw;
    const size_t wn = (size_t)p.fs * p.Cin * p.Cout;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dgd, gdec.size() * 4)); CK(cudaMalloc(&dgo, godd.size() * 4)); CK(cudaMalloc(&ddw, wn * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dgd, gdec.data(), gdec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dgo, godd.data(), godd.size() * 4, cudaMemcpyHostToDevice));
    UmmaWgradLaunch WL;
    memset(&WL, 0, sizeof(WL));
    WL.batch = p.B; WL.dW = ddw; WL.w_sp = p.Cout; WL.w_sg = 1; WL.scale = 1.f;
    for (int q = 0; q < 2; ++q)
        for (int par = 0; par < 2; ++par) {
            WgGroup W;
            memset(&W, 0, sizeof(W));
            W.P.base = dx + par * p.Cin; W.P.bstride = (long long)p.T * p.Cin; W.P.rstride = 2 * p.Cin;
            W.P.r_lo = 0; W.P.r_hi = (par == 0) ? (p.T + 1) / 2 : p.T / 2; W.P.C = p.Cin; W.P.kind = PLANE_DIRECT;
            W.G.base = (q == 0) ? dgd : dgo - (long long)mo_lo * p.Cout;
            W.G.bstride = (q == 0) ? (long long)Td * p.Cout : (long long)n_odd * p.Cout;
            W.G.rstride = p.Cout; W.G.C = p.Cout; W.G.kind = PLANE_DIRECT;
            W.m_lo = (q == 0) ? 0 : mo_lo; W.m_hi = (q == 0) ? Td : mo_hi;
            W.G.r_lo = W.m_lo; W.G.r_hi = W.m_hi;
            for (int j = 0; j < p.fs; ++j) {
                int e = q + j;
                if ((e & 1) != par) continue;
                W.d[W.ntaps] = e >> 1; W.woff[W.ntaps] = j * p.Cin * p.Cout; ++W.ntaps;
            }
            if (W.m_hi <= W.m_lo || W.ntaps == 0) continue;
            WL.grp[WL.ngroups++] = W;
        }
    if (!umma_plan_wgrad(&WL)) { printf("[%s] wgrad not eligible\n", name); exit(4); }
    const WgGroup& W0 = WL.grp[0];
    printf("[%s] wgrad B=%d T=%d Cin=%d Cout=%d swap=%d NT=%d taps/cta=%d tapsets=%d chunks/cta=%d grid=(%d,%d,%d) stages=%d smem=%zu\n", name,
           p.B, p.T, p.Cin, p.Cout, W0.swap, W0.NT, W0.taps_per_cta, W0.n_tapsets, W0.chunks_per_cta, WL.grid_x, WL.grid_y, WL.grid_z,
           WL.nstages, umma_wgrad_smem_bytes(WL));

    // split arena exactly as the engine builds it: one array per distinct plane view, one batched split pass
    const size_t arena_bytes = umma_plan_wgrad_split(WL, p.B, nullptr, nullptr, nullptr);
    if (!arena_bytes) { printf("[%s] too many distinct views for one split pass\n", name); return; }
    uint8_t* arena;
    CK(cudaMalloc(&arena, arena_bytes));
    WgSplit S;
    SplitJobs J;
    umma_plan_wgrad_split(WL, p.B, arena, &S, &J);
    printf("[%s] split arena %.1f MB in %d arrays\n", name, arena_bytes / 1e6, J.njobs);
    CK(launch_split_views(J, 0));
    CK(cudaDeviceSynchronize());

    auto check_dw = [&](const char* which) {
        std::vector<float> dw(wn);
        CK(cudaMemcpy(dw.data(), ddw, wn * 4, cudaMemcpyDeviceToHost));
        std::vector<double> ref(wn, 0.0);
        auto split = [](float v, float* h, float* l) { *h = bf16_round(v); *l = bf16_round(v - *h); };
        for (int b = 0; b < p.B; ++b)
            for (int a = 0; a < To; ++a) {
                const bool even = (a & 1) == 0;
                if (!even && !(a >= p.cs && a < p.cs + p.U)) continue;
                const float* g = even ? &gdec[((size_t)b * Td + a / 2) * p.Cout] : &godd[((size_t)b * n_odd + ((a - 1) / 2 - mo_lo)) * p.Cout];
                for (int j = 0; j < p.fs; ++j)
                    for (int c = 0; c < p.Cin; ++c) {
                        float xh, xl; split(x[((size_t)b * p.T + a + j) * p.Cin + c], &xh, &xl);
                        for (int n = 0; n < p.Cout; ++n) {
                            float gh, gl; split(g[n], &gh, &gl);
                            ref[((size_t)j * p.Cin + c) * p.Cout + n] += (double)xh * gh + (double)xl * gh + (double)xh * gl;
                        }
                    }
            }
        double num = 0, den = 0;
        for (size_t i = 0; i < wn; ++i) { num += (dw[i] - ref[i]) * (dw[i] - ref[i]); den += ref[i] * ref[i]; }
        const double rel = sqrt(num / den);
        printf("[%s] wgrad %-9s %s rel_l2=%.3e\n", name, which, rel < 1e-4 ? "PASS" : "FAIL", rel);
    };
    auto time_it = [&](const char* which, auto launch) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        launch();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < timing_iters; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= timing_iters;
        const double flops = 2.0 * p.B * ((double)Td + n_odd) * p.fs * p.Cin * p.Cout;
        printf("[%s] wgrad %-9s time %.1f us  %.1f useful TFLOP/s\n", name, which, ms * 1e3, flops / (ms * 1e-3) * 1e-12);
    };
    CK(cudaMemset(ddw, 0, wn * 4));
    CK(launch_wgrad_umma(WL, 0));
    { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("[%s] engine WGRAD ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (check) check_dw("engine");
    CK(cudaMemset(ddw, 0, wn * 4));
    CK(launch_wgrad_umma_bulk(WL, S, 0));
    { cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[%s] bulk-fed WGRAD ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (check) check_dw("bulk-fed");
    if (timing_iters > 0) {
        time_it("engine", [&]() { CK(launch_wgrad_umma(WL, 0)); });
        time_it("bulk-fed", [&]() { CK(launch_wgrad_umma_bulk(WL, S, 0)); });
        time_it("split pass", [&]() { CK(launch_split_views(J, 0)); });
    }
    cudaFree(dx); cudaFree(dgd); cudaFree(dgo); cudaFree(ddw); cudaFree(arena);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    const char* which = argc > 1 ? argv[1] : "all";
    auto want = [&](const char* n) { return !strcmp(which, "all") || !strcmp(which, n); };
    if (want("small")) {
        run_case("tiny",    { 1, 300,  16, 16,  3,  40, 101}, true, 0);
        run_case("taps15",  { 2, 1500, 32, 48, 15, 200, 401}, true, 0);
        run_case("c72_n96", { 2, 3000, 72, 96, 15, 500, 801}, true, 0);
        run_case("c24",     { 2, 1500, 24, 48, 15, 200, 401}, true, 0);
    }
    // M4 down blocks at B=16 (T = input rows of the layer)
    if (want("down1")) run_case("down1", {16, 73715, 24, 48, 15, 32750, 8201}, true, 20);
    if (want("down2")) run_case("down2", {16, 36851, 48, 72, 15, 16366, 4105}, false, 20);
    if (want("down3")) run_case("down3", {16, 18419, 72, 96, 15, 8174, 2057}, false, 20);
    if (want("down4")) run_case("down4", {16, 9203, 96, 120, 15, 4078, 1033}, false, 20);
    // deep (sparse, fewer tiles than SMs) layers: is the converter the limiter there?
    if (want("down5")) run_case("down5", {16, 4595, 120, 144, 15, 2030, 521}, false, 20);
    if (want("down7")) run_case("down7", {16, 1139, 168, 192, 15, 494, 137}, false, 20);
    if (want("down9")) run_case("down9", {16, 275, 216, 240, 15, 110, 41}, false, 20);
    if (want("down11")) run_case("down11", {16, 59, 264, 288, 15, 14, 17}, false, 20);
    if (want("wsmall")) {
        run_wgrad("wg_tiny",   { 1, 300,  16, 16,  3,  40, 101}, true, 0);
        run_wgrad("wg_taps15", { 2, 1500, 32, 48, 15, 200, 401}, true, 0);
        run_wgrad("wg_c72n96", { 2, 3000, 72, 96, 15, 500, 801}, true, 0);
        run_wgrad("wg_c24",    { 2, 1500, 24, 48, 15, 200, 401}, true, 0);
        run_wgrad("wg_wide",   { 2, 300, 264, 288, 15, 50, 101}, true, 0);
    }
    if (want("wdown1")) run_wgrad("wg_down1", {16, 73715, 24, 48, 15, 32750, 8201}, false, 10);
    if (want("wdown2")) run_wgrad("wg_down2", {16, 36851, 48, 72, 15, 16366, 4105}, false, 10);
    if (want("wdown3")) run_wgrad("wg_down3", {16, 18419, 72, 96, 15, 8174, 2057}, false, 10);
    if (want("wdown5")) run_wgrad("wg_down5", {16, 4595, 120, 144, 15, 2030, 521}, false, 10);
    return 0;
}
