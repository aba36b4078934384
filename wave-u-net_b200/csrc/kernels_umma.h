// kernels_umma.h - parameter blocks and launchers of the tcgen05 plane-convolution kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "launch.h"

namespace wun {

constexpr int kUmmaMaxGroups = 4;
constexpr int kUmmaMaxSplit = 4;

struct UmmaGroup {          // all terms of one class that read the same plane (consecutive row shifts)
    int plane;
    int term_begin, term_end;   // indices into UmmaLaunch::d
    int dmin;
};

struct UmmaClass {
    OutView out;
    UmmaGroup groups[kUmmaMaxGroups];
    int ngroups;
    const uint8_t* wpack[kUmmaMaxSplit];   // packed weights of this class, per output-channel split
};

struct UmmaLaunch {
    PlaneView planes[kMaxPlanes];
    UmmaClass cls[kMaxClasses];
    int d[kMaxTerms];       // row shift of every term (grouped per class/plane)
    int ncls;
    int N;                  // real output channels
    int NPAD;               // channels handled per split, multiple of 16, <= 256
    int nsplit;             // output-channel splits (N > 256, or to spread sparse launches over more SMs)
    int MT;                 // 128-row tiles per CTA
    int rows_alloc;         // slab rows per stage  (>= MT*128 + max row-shift span)
    int tmem_cols;          // power of two >= MT*NPAD
    int TB, nbs;            // weight ring: nbs stages of TB taps
    int nteams;             // converter teams of the non-persistent kernel: 2 (dense launches) or 4 (sparse)
    int persistent;         // 1: one-CTA-per-SM tile loop with double-buffered TMEM (big layers)
    int fuse;               // 1: fused-N MMAs over [B_hi | B_lo] (2 per product; accumulator tile = 2*NPAD columns)
    int consec;             // 1: within every group the terms are sorted by row shift, d = dmin, dmin+1, ... (issue_taps path)
    int folded;             // 1: batch-folded row tiles + cluster split-K (plane_conv_umma_fold; the deep, few-row layers)
    int fold_flags;         // folded: bit 0 = one-row converter passes (compact code), bit 1 = L2 prefetch of the CTA's weights
    int ksplit;             // folded: CTAs per cluster; each takes a contiguous range of the tile's (plane, chunk) jobs
    int fold_pitch[kMaxClasses];   // folded: virtual rows per batch item of a class = its rows + its widest tap span
    const float* bias;
    int epilogue;
    int batch;
    int pairC;              // > 0: pair-merged classes (launch.h OutView::pairC); the bias repeats with this period
    int epi2;               // persistent dgrad: two epilogue warp groups (and two converter teams) - plane_conv_umma_persistent_dg2
    int epi_teams;          // 2-CTA / sparse kernel: converter teams that share the epilogue (2; WUN_EPI_TEAMS=1 = team 0 alone)
};

constexpr int kUmmaMaxPackJobs = kMaxClasses * kUmmaMaxSplit;

struct UmmaPackJob {        // one (class, split) weight pack
    uint8_t* out;
    int ngroups;
    int g_nchunk[kUmmaMaxGroups], g_nterm[kUmmaMaxGroups], g_term_begin[kUmmaMaxGroups], g_C[kUmmaMaxGroups];
    int n0;
    int nblocks;
};

struct UmmaPackLaunch {
    const float* W;
    int woff[kMaxTerms];
    int woff2[kMaxTerms];   // pair-merged launches (pairC > 0): weight offset of columns [pairC, 2*pairC); -1 = zeros
    int pairC;
    int w_sk, w_sn;
    int N, NPAD;
    int njobs;
    UmmaPackJob jobs[kUmmaMaxPackJobs];
};

struct UmmaChoice {         // tiling decisions for one ConvLaunch
    int NPAD, nsplit, MT, rows_alloc, tmem_cols, TB, nbs, persistent, nteams, fuse;
    int folded, ksplit;     // batch-folded cluster split-K kernel (sparse launches)
    int epi2;               // see UmmaLaunch
    size_t pack_bytes;      // arena bytes the packed weights of this launch need
};

// wgrad on tensor cores: for every group (one forward class x one input plane)
//   dW[woff_t + cp*w_sp + cg*w_sg] += scale * sum_{b, m in [m_lo,m_hi)} P[b, m+d_t, cp] * G[b, m, cg]
// A GEMM with the ROWS as the reduction dimension: both operands are MN-major (channels contiguous).
constexpr int kWgMaxGroups = 8;
constexpr int kWgMaxTaps = 8;

struct WgGroup {
    PlaneView P;            // activation side (tap-shifted; may be a MID plane)
    PlaneView G;            // gradient side (pre-activation gradient), rows m
    int m_lo, m_hi;
    int ntaps;
    int d[kWgMaxTaps];
    int woff[kWgMaxTaps];
    // tiling (filled by umma_plan_wgrad)
    int swap;               // 0: A (M side) = P, B (N side) = G;  1: A = G, B = P
    int NT;                 // N-tile width (multiple of 16, <= 128)
    int n_mtiles, n_ntiles; // tiles of 128 / NT channels
    int taps_per_cta, n_tapsets;
    int chunks_per_batch;   // 64-row chunks per batch element
    int chunks_per_cta;     // consecutive (batch-folded) chunks one CTA reduces
    int n_ctas_x;           // ceil(batch*chunks_per_batch / chunks_per_cta)
    int tmem_cols;
    int fuse;               // bulk-fed kernel: fused-N MMAs (A_hi x [B_hi|B_lo], A_lo x B_hi); accumulator stride 2*NT per tap
    int z0;                 // first blockIdx.z of this group (one z per tap set)
};

struct UmmaWgradLaunch {
    WgGroup grp[kWgMaxGroups];
    int ngroups;
    int batch;
    float* dW;
    int w_sp, w_sg;         // element strides of dW along the P-channel / G-channel index
    float scale;
    int grid_x, grid_y, grid_z;
    int nstages;            // shared-memory pipeline stages (2 or 3)
};

cudaError_t launch_wgrad_umma(const UmmaWgradLaunch& L, cudaStream_t stream);
// ---- bulk-copy-fed wgrad (experimental, WUN_BULK_WGRAD=1) --------------------------------------------------------
constexpr int kSplitMaxJobs = 2 * kWgMaxGroups;

struct SplitJob {           // materialise plane rows [row0, row0 + rows) of V as hi/lo bf16 atom planes at `out`
    PlaneView V;
    uint8_t* out;           // [batch][nchunk][4][rows][16 B]
    int nchunk, rows, row0;
    float* colsum;          // class-gradient views: bias gradient [V.C] that receives colsum_scale * column sums (else null)
};

struct SplitJobs {
    SplitJob job[kSplitMaxJobs];
    int njobs, batch;
    float colsum_scale;
};

struct WgSplit {            // where the wgrad groups find their operands (filled next to the SplitJobs)
    const uint8_t* P[kWgMaxGroups];
    const uint8_t* G[kWgMaxGroups];
    long long p_pstride[kWgMaxGroups], g_pstride[kWgMaxGroups];   // bytes between the 4 sub-planes of a chunk (= rows * 16)
    int p_nchunk[kWgMaxGroups], g_nchunk[kWgMaxGroups];
    int p_row0[kWgMaxGroups], g_row0[kWgMaxGroups];               // plane row stored at array index 0
};

// Split arena of one layer's wgrad at `batch`: bytes needed (0 = too many distinct views); fills the jobs / operand table
// when `arena` (256-B aligned, that many bytes) is given.
size_t umma_plan_wgrad_split(const UmmaWgradLaunch& U, int batch, uint8_t* arena, WgSplit* S, SplitJobs* J);
cudaError_t launch_split_views(const SplitJobs& J, cudaStream_t stream);
cudaError_t launch_wgrad_umma_bulk(const UmmaWgradLaunch& L, const WgSplit& S, cudaStream_t stream);
// rows one wgrad CTA can read past the last valid row of a group: G side / P side (chunk rounding + tap span)
constexpr int kWgOverreachG = 64, kWgOverreachP = 64 + 16;

// fills the tiling fields (groups' P, G, taps and the common fields already set); false = not eligible
bool umma_plan_wgrad(UmmaWgradLaunch* L);

// Output layer fused into the epilogue of the last up block's forward conv (persistent kernel only; north_star "the per-source
// tanh / difference OutputLayer is fused into the final conv").  The conv epilogue holds one feature row per thread: it applies the
// 1x1 output convs over [crop(mix) || features] (OutputLayer.py:8,15), the activation, the difference source (:17-22) and the test-time
// clip, writes the source estimates, and - when targets are given - the MSE loss (Training.py:50-63), dL/dpre of the output
// convs, their weight / bias gradients and the gradient w.r.t. the features' pre-activation, so that output_fwd / output_dgrad /
// output_wgrad launch nothing.
constexpr int kOutFuseMaxCols = 8;    // nconv * C the fused epilogue supports (registers); wider heads use the separate kernels
struct OutputFuse {
    OutputLaunch O;
    float* gfeat;           // training: dPre of the last up block (gradient twin of O.feat, same geometry) or null
    float* grads;           // training: flat gradient buffer (O.w_off / O.b_off index it) or null
    float grad_scale;
    int t_step;             // output frame of (class q, row m, half h) = t0[q] + t_step * m + h
    int t0[kMaxClasses];
};
// whether the conv launch `ch` plans for L can carry the output layer described by O
bool umma_output_fusable(const ConvLaunch& L, const UmmaChoice& ch, const OutputLaunch& O);

size_t umma_smem_bytes(const UmmaLaunch& L);
// dynamic shared memory the kernel variant chosen in `ch` is launched with / of a planned wgrad launch (plan audit)
size_t umma_choice_smem_bytes(const UmmaChoice& ch);
size_t umma_wgrad_smem_bytes(const UmmaWgradLaunch& L);
cudaError_t launch_plane_conv_umma(const UmmaLaunch& L, cudaStream_t stream, const OutputFuse* fuse = nullptr);
cudaError_t launch_umma_pack(const UmmaPackLaunch& PL, cudaStream_t stream);
// Decide whether / how a generic plane-convolution launch runs on tcgen05; false = not eligible (SIMT).
bool umma_plan_from_conv(const ConvLaunch& L, UmmaChoice* choice);
// Pack the weights into `arena` (choice.pack_bytes, 256-B aligned) and enqueue the tcgen05 kernel.
cudaError_t umma_run_conv(const ConvLaunch& L, const UmmaChoice& choice, uint8_t* arena, cudaStream_t stream);
// The two parameter blocks umma_run_conv launches (exposed for the probe / per-kernel timing).
cudaError_t umma_build(const ConvLaunch& L, const UmmaChoice& choice, uint8_t* arena, UmmaLaunch* U, UmmaPackLaunch* PL);

}  // namespace wun
