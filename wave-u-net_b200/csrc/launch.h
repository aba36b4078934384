// launch.h - by-value kernel parameter blocks shared by the host planner and the CUDA kernels.
//
// Vocabulary (DESIGN.md "Plane convolution"):
//   plane  : a strided view  (b, r, c) -> T[b][row_offset + r*row_step][c]  of an fp32 [B,rows,C]
//            tensor, valid for r in [r_lo, r_hi) and ZERO outside (this is how "same" padding, the
//            skip crop, decimation parity and dgrad boundaries are all expressed).  A MID plane is
//            computed on load from two neighbouring rows of its tensor (linear / learned upsampling,
//            reference UnetAudioSeparator.py:109-118, InterpolationLayer.py:19-39).
//   class  : one family of output rows  m in [m_lo, m_hi)  written through an output view.
//   term   : one (plane, row shift d, weight slice) contribution  plane[b, m+d, :] @ W_term.
// Every 1-D convolution of the network - forward, its dgrad, and (with rows as the reduction axis) its
// wgrad - is a sum of terms over planes.
#pragma once
#include <stdint.h>

#define WUN_MAX_OUT_COLS 16   // nconv * num_channels of the output layer (<= 8 sources * 2 channels)

namespace wun {

constexpr int kMaxPlanes = 4;
constexpr int kMaxClasses = 4;
constexpr int kMaxTerms = 72;

enum PlaneKind : int { PLANE_DIRECT = 0, PLANE_MID = 1 };
// boundary handling of the "next" row of a MID plane whose tensor has xrows rows
enum MidMode : int { MID_VALID = 0,   // r+1 always exists (context / 'valid')
                     MID_CLAMP = 1,   // next = x[min(r+1, N-1)]  (legacy resize_bilinear, non-context)
                     MID_ZERO = 2 };  // next = 0 for r+1 == N    (learned interpolation, 'same')

struct PlaneView {
    const float* base;     // tensor base + row_offset*C (may point before the allocation; guarded)
    long long bstride;     // elements between batch items
    int rstride;           // elements between consecutive plane rows (row_step*C)
    int r_lo, r_hi;        // valid plane rows
    int C;                 // channels of the plane
    int kind;              // PlaneKind
    int mid_mode;          // MidMode
    int xrows;             // rows of the underlying tensor (MID: index of last row = xrows-1)
    const float* blend;    // MID: per-channel weight w (sigmoid(var)) or nullptr for 0.5
};

struct OutView {
    float* base;           // tensor base + row_offset*C
    long long bstride;
    int rstride;
    int m_lo, m_hi;        // rows this class produces
    const float* saved;    // dgrad: forward activation with the same geometry (LeakyReLU slope) or null
    int acc_lo, acc_hi;    // rows in [acc_lo, acc_hi) accumulate into the destination (+=)
    int term_begin, term_end;
    // Pair-merged class (pairC > 0; dgrad only): the launch computes N = 2*pairC columns per row m.  Columns [0, pairC) are
    // the class above restricted to rows [lo_h[0], hi_h[0]); columns [pairC, 2*pairC) are a SECOND class with its own view,
    // rows [lo_h[1], hi_h[1]) - e.g. the gradients of the even and the odd input rows of a down block (one tensor viewed as
    // [B, rows/2, 2C]: base2 = base + C, same strides), or of the skip-even / skip-odd and copied / interpolated inputs of an
    // up block.  Both classes read the same planes with the same row shifts, so every slab is staged (and every tap issued)
    // once for both; m_lo / m_hi are the union of the two row ranges.
    int pairC;
    float* base2;
    long long bstride2;
    int rstride2;
    const float* saved2;
    int acc_lo2, acc_hi2;
    int lo_h[2], hi_h[2];
};

struct Term {
    int plane;             // index into planes[]
    int d;                 // row shift: reads plane row m + d
    int woff;              // element offset of W[tap][coff][0]; pair-merged classes: -1 = this half has no such tap (zeros)
    int woff2;             // pair-merged classes: the same for columns [pairC, 2*pairC)
};

enum Epilogue : int { EPI_BIAS_LRELU = 0,   // y = leaky_relu(acc + bias)          (forward)
                      EPI_SLOPE = 1,        // y = acc * slope(saved) [+ old]      (dgrad)
                      EPI_PLAIN = 2 };      // y = acc [+ old]

struct ConvLaunch {
    PlaneView planes[kMaxPlanes];
    OutView cls[kMaxClasses];
    Term terms[kMaxTerms];
    int nplanes, ncls;
    int N;                 // output channels of this launch
    int w_sk, w_sn;        // weight element strides along the reduction channel / output channel
    const float* W;
    const float* bias;     // may be null
    int epilogue;
    int batch;
    int max_rows;          // max over classes of (m_hi - m_lo)
    int pairC;             // > 0: every class of this launch is pair-merged with this half width (N = 2*pairC)
};

// wgrad: dW[woff + c*w_sk + n*w_sn] += sum_{b, m in [m_lo,m_hi)} plane[b, m+d, c] * dpre[b, m, n]
struct WgradLaunch {
    PlaneView plane;       // activation side (may be MID)
    PlaneView dpre;        // gradient side, always DIRECT; row m
    int m_lo, m_hi;
    int nterms;
    int d[16];
    int woff[16];
    int N;                 // channels of dpre
    int w_sk, w_sn;
    float* dW;
    float scale;
    int batch;
    int rows_per_cta;
};

}  // namespace wun
