// plan.h - host-side description of the network as plane convolutions (no CUDA types).
//
// build_plan() restates the shape algebra of the reference
//   get_padding            /root/reference/Models/UnetAudioSeparator.py:34-83
//   get_output             /root/reference/Models/UnetAudioSeparator.py:85-144
//   crop / crop_and_concat /root/reference/Utils.py:11-24,104-123
// and records, for every conv layer, which rows are LIVE (survive the [:, ::2, :] decimation at :100 or
// the centre crop of the skip connection at :122) and how each live output row reads its inputs.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/wun.h"

namespace wun {

constexpr int TENSOR_MIX = -1;   // the caller's input batch

struct TensorSpec {
    std::string name;
    int64_t rows = 0;      // per batch element
    int C = 0;
    bool per_batch = true; // false: size rows*C independent of the batch (sigmoid(var) vectors)
    bool training_only = false;
};

struct ViewSpec {
    int tensor = 0;
    int row_offset = 0, row_step = 1;   // element (b, r, c) = T[b][row_offset + r*row_step][c]
    int r_lo = 0, r_hi = 0;             // valid rows (zero outside)
    int C = 0;
    int kind = 0, mid_mode = 0, xrows = 0;
    int blend_tensor = -2;              // tensor holding sigmoid(var) (learned) or -2
};

struct TermSpec {
    int plane, d, tap, coff;
};

struct ClassSpec {
    ViewSpec out;                        // r == m
    int m_lo = 0, m_hi = 0;
    std::vector<TermSpec> terms;         // sorted by (plane, d)
};

struct ConvOp {
    std::string name;
    int k = 0, cin_tot = 0, cout = 0;
    int w_param = -1, b_param = -1;      // indices into the parameter table
    std::vector<ViewSpec> planes;
    std::vector<int> plane_grad_tensor;  // where d(plane) goes: a tensor id, or -2 = not needed
    std::vector<int> plane_slope;        // 1: multiply by LeakyReLU slope of the plane's own tensor
    std::vector<ClassSpec> classes;
};

struct UpsampleSpec {                    // one per up level
    int src_tensor;                      // tensor that is upsampled (z or up[i-1])
    int N, nmid, C, mid_mode;
    int interp_param;                    // parameter index of interp_<level> or -1
    int wsig_tensor;                     // tensor id of sigmoid(var) or -2
};

struct Plan {
    WunConfig cfg;
    int64_t T_in = 0, T_out = 0, Tf = 0;
    int crop_feat = 0, crop_out = 0, out_pad_left = 0, nconv = 0;
    std::vector<WunParamInfo> params;
    int64_t param_numel = 0;
    std::vector<int> out_w_param, out_b_param;
    std::vector<TensorSpec> tensors;
    std::vector<int> grad_twin;          // tensor id -> id of its gradient tensor (or -2)
    std::vector<ConvOp> down, up;        // L each
    ConvOp bottleneck;
    std::vector<UpsampleSpec> ups;
    int t_gue = -2, t_gmid = -2, t_dpre_out = -2, t_outbuf = -2, t_feat = -2;
    double fwd_flops_per_item = 0, dgrad0_flops_per_item = 0;
    std::string describe() const;
};

// returns WUN_OK or an error code; msg receives the reason
int solve_padding(const WunConfig& cfg, int64_t num_frames, int64_t* t_in, int64_t* t_out, std::string* msg);
// t_in_frames: input window length; every other length follows from it (get_output :97-127)
int build_plan(const WunConfig& cfg, int64_t t_in_frames, Plan* plan, std::string* msg);

}  // namespace wun
