// plan.cpp - shape solver and plane-convolution planner (pure host code).
#include "plan.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "launch.h"

namespace wun {

static inline int floordiv2(int e) { return (e >= 0) ? (e / 2) : -((-e + 1) / 2); }
static inline int mod2(int e) { return ((e % 2) + 2) % 2; }

int solve_padding(const WunConfig& c, int64_t num_frames, int64_t* t_in, int64_t* t_out, std::string* msg) {
    // UnetAudioSeparator.py:34-83
    if (!c.context) {
        *t_in = num_frames;
        *t_out = num_frames;
        return WUN_OK;
    }
    double rem = (double)num_frames;
    rem = rem - c.output_filter_size + 1;                         // :46
    for (int i = 0; i < c.num_layers; ++i) {                      // :49-51
        rem = rem + c.merge_filter_size - 1;
        rem = (rem + 1.0) / 2.0;
    }
    int64_t x = (int64_t)ceil(rem);                               // :54
    if (x < 2) {                                                  // :55
        if (msg) *msg = "get_padding: bottleneck feature map narrower than 2 frames (reference assert x >= 2)";
        return WUN_E_SHAPE;
    }
    int64_t out = x, in = x + c.filter_size - 1;                  // :58-62
    for (int i = 0; i < c.num_layers; ++i) {                      // :65-73
        out = 2 * out - 1;
        out = out - c.merge_filter_size + 1;
        in = 2 * in - 1;
        in = in + ((i < c.num_layers - 1) ? c.filter_size : c.input_filter_size) - 1;
    }
    out = out - c.output_filter_size + 1;                         // :76
    *t_in = in;
    *t_out = out;
    return WUN_OK;
}

static int add_tensor(Plan* p, const std::string& name, int64_t rows, int C, bool per_batch, bool training_only) {
    TensorSpec t;
    t.name = name; t.rows = rows; t.C = C; t.per_batch = per_batch; t.training_only = training_only;
    p->tensors.push_back(t);
    p->grad_twin.push_back(-2);
    return (int)p->tensors.size() - 1;
}

static int add_act(Plan* p, const std::string& name, int64_t rows, int C) {
    int t = add_tensor(p, name, rows, C, true, false);
    int g = add_tensor(p, "g_" + name, rows, C, true, true);
    p->grad_twin[t] = g;
    return t;
}

static void sort_terms(ClassSpec* c) {
    std::stable_sort(c->terms.begin(), c->terms.end(), [](const TermSpec& a, const TermSpec& b) {
        if (a.plane != b.plane) return a.plane < b.plane;
        return a.d < b.d;
    });
}

static ViewSpec direct_view(int tensor, int row_offset, int row_step, int r_lo, int r_hi, int C) {
    ViewSpec v;
    v.tensor = tensor; v.row_offset = row_offset; v.row_step = row_step; v.r_lo = r_lo; v.r_hi = r_hi; v.C = C;
    v.kind = PLANE_DIRECT;
    return v;
}

static double op_flops(const ConvOp& op) {
    double f = 0;
    for (const auto& c : op.classes) {
        double ksum = 0;
        for (const auto& t : c.terms) ksum += op.planes[t.plane].C;
        f += 2.0 * (double)std::max(0, c.m_hi - c.m_lo) * ksum * op.cout;
    }
    return f;
}

int build_plan(const WunConfig& cfg, int64_t t_in_frames, Plan* P, std::string* msg) {
    auto fail = [&](int code, const char* m) { if (msg) *msg = m; return code; };
    if (cfg.num_layers < 1 || cfg.num_layers > 24) return fail(WUN_E_INVALID, "num_layers out of range [1,24]");
    if (cfg.num_initial_filters < 1) return fail(WUN_E_INVALID, "num_initial_filters < 1");
    if (cfg.filter_size < 1 || cfg.merge_filter_size < 1 || cfg.output_filter_size < 1 || cfg.input_filter_size < 1)
        return fail(WUN_E_INVALID, "filter sizes must be >= 1");
    if (cfg.filter_size > 63 || cfg.merge_filter_size > 31) return fail(WUN_E_INVALID, "filter size too large");
    if (cfg.upsampling != 0 && cfg.upsampling != 1) return fail(WUN_E_NOTIMPL, "upsampling must be linear or learned");
    if (cfg.output_type != 0 && cfg.output_type != 1)
        return fail(WUN_E_NOTIMPL, "output_type must be direct or difference (UnetAudioSeparator.py:144)");
    if (cfg.output_activation != 0 && cfg.output_activation != 1)
        return fail(WUN_E_NOTIMPL, "output_activation must be tanh or linear (UnetAudioSeparator.py:136)");
    if (cfg.num_channels < 1 || cfg.num_channels > 2) return fail(WUN_E_INVALID, "num_channels must be 1 or 2");
    if (cfg.num_sources < 1 || cfg.num_sources > WUN_MAX_SOURCES) return fail(WUN_E_INVALID, "num_sources out of range");
    if (cfg.output_type == 1 && cfg.num_sources < 2) return fail(WUN_E_INVALID, "difference output needs >= 2 sources");

    *P = Plan();
    P->cfg = cfg;
    P->T_in = t_in_frames;
    if (P->T_in > (1 << 30) || P->T_in < 1) return fail(WUN_E_SHAPE, "window too long or empty");

    const int L = cfg.num_layers, F0 = cfg.num_initial_filters, C = cfg.num_channels;
    const int fs = cfg.filter_size, mfs = cfg.merge_filter_size, ofs = cfg.output_filter_size;
    const bool ctx = cfg.context != 0, learned = cfg.upsampling == 1;
    P->nconv = (cfg.output_type == 0) ? cfg.num_sources : cfg.num_sources - 1;
    if (P->nconv * C > WUN_MAX_OUT_COLS) return fail(WUN_E_INVALID, "too many output columns");

    // ---- shapes (get_output :97-127) --------------------------------------------------------
    std::vector<int> T(L + 1), To(L), Td(L), Nup(L + 1), U(L), nmid(L), cs(L), mo_lo(L), mo_hi(L);
    T[0] = (int)P->T_in;
    for (int i = 0; i < L; ++i) {
        To[i] = ctx ? T[i] - fs + 1 : T[i];
        if (To[i] < 1) return fail(WUN_E_SHAPE, "input too short for the down-sampling path");
        Td[i] = (To[i] + 1) / 2;                                  // [:, ::2, :]  (:100)
        T[i + 1] = Td[i];
    }
    const int Tb = ctx ? T[L] - fs + 1 : T[L];                    // :102
    if (Tb < 1) return fail(WUN_E_SHAPE, "input too short for the bottleneck conv");
    Nup[0] = Tb;
    for (int i = 0; i < L; ++i) {
        const int j = L - 1 - i;
        U[i] = ctx ? 2 * Nup[i] - 1 : 2 * Nup[i];                 // :115 / :117 / InterpolationLayer.py:30
        nmid[i] = ctx ? Nup[i] - 1 : Nup[i];
        if (!ctx && To[j] != U[i])                                // :121
            return fail(WUN_E_SHAPE, "skip and upsampled lengths differ without context (reference assert :121); "
                                     "num_frames must be divisible by 2^num_layers");
        if (To[j] < U[i]) return fail(WUN_E_SHAPE, "crop: negative difference (Utils.py:117)");
        cs[i] = (To[j] - U[i]) / 2;                               // Utils.py:120
        const int V = ctx ? U[i] - mfs + 1 : U[i];                // :123
        if (V < 1) return fail(WUN_E_SHAPE, "merge conv output empty");
        Nup[i + 1] = V;
        mo_lo[j] = cs[i] / 2;
        mo_hi[j] = (cs[i] + U[i]) / 2;
    }
    P->Tf = Nup[L];
    if (P->T_in < P->Tf) return fail(WUN_E_SHAPE, "crop: negative difference (Utils.py:117)");
    P->crop_feat = (int)((P->T_in - P->Tf) / 2);                  // :127
    const int64_t tout = ctx ? P->Tf - ofs + 1 : P->Tf;
    if (tout < 1) return fail(WUN_E_SHAPE, "output conv output empty");
    P->T_out = tout;                                              // what get_output really produces
    P->crop_out = (int)((P->T_in - P->T_out) / 2);                // OutputLayer.py:20
    P->out_pad_left = ctx ? 0 : (ofs - 1) / 2;

    // ---- parameter table (TF creation order) -----------------------------------------------------
    int nconv_layers = 0;
    auto add_param = [&](const char* nm, int ndim, int s0, int s1, int s2) {
        WunParamInfo pi;
        memset(&pi, 0, sizeof(pi));
        snprintf(pi.name, sizeof(pi.name), "%s", nm);
        pi.ndim = ndim; pi.shape[0] = s0; pi.shape[1] = s1; pi.shape[2] = s2;
        pi.offset = P->param_numel;
        pi.numel = (int64_t)s0 * (ndim > 1 ? s1 : 1) * (ndim > 2 ? s2 : 1);
        P->param_numel += pi.numel;
        P->params.push_back(pi);
        return (int)P->params.size() - 1;
    };
    auto add_conv_params = [&](int k, int cin, int cout, int* w, int* b) {
        char nm[64], sfx[16] = "";
        if (nconv_layers > 0) snprintf(sfx, sizeof(sfx), "_%d", nconv_layers);
        snprintf(nm, sizeof(nm), "separator/conv1d%s/kernel", sfx);
        *w = add_param(nm, 3, k, cin, cout);
        snprintf(nm, sizeof(nm), "separator/conv1d%s/bias", sfx);
        *b = add_param(nm, 1, cout, 1, 1);
        ++nconv_layers;
    };

    // ---- tensors ------------------------------------------------------------------------------
    std::vector<int> t_dec(L), t_odd(L), t_up(L);
    for (int i = 0; i < L; ++i) {
        t_dec[i] = add_act(P, "dec" + std::to_string(i), Td[i], F0 * (i + 1));
        t_odd[i] = add_act(P, "odd" + std::to_string(i), std::max(0, mo_hi[i] - mo_lo[i]), F0 * (i + 1));
    }
    const int t_z = add_act(P, "z", Tb, F0 * (L + 1));
    for (int i = 0; i < L; ++i) t_up[i] = add_act(P, "up" + std::to_string(i), Nup[i + 1], F0 * (L - i));
    P->t_feat = t_up[L - 1];

    // ---- down path ------------------------------------------------------------------------------
    P->down.resize(L);
    for (int i = 0; i < L; ++i) {
        ConvOp& op = P->down[i];
        op.name = "down" + std::to_string(i);
        op.k = fs; op.cin_tot = (i == 0) ? C : F0 * i; op.cout = F0 * (i + 1);
        add_conv_params(fs, op.cin_tot, op.cout, &op.w_param, &op.b_param);
        const int x = (i == 0) ? TENSOR_MIX : t_dec[i - 1];
        const int padL = ctx ? 0 : (fs - 1) / 2;
        op.planes.push_back(direct_view(x, 0, 2, 0, (T[i] + 1) / 2, op.cin_tot));
        op.planes.push_back(direct_view(x, 1, 2, 0, T[i] / 2, op.cin_tot));
        const int gx = (i == 0) ? -2 : P->grad_twin[x];
        op.plane_grad_tensor = {gx, gx};
        op.plane_slope = {1, 1};
        ClassSpec c0, c1;
        c0.out = direct_view(t_dec[i], 0, 1, 0, Td[i], op.cout);
        c0.m_lo = 0; c0.m_hi = Td[i];
        c1.out = direct_view(t_odd[i], -mo_lo[i], 1, mo_lo[i], mo_hi[i], op.cout);
        c1.m_lo = mo_lo[i]; c1.m_hi = std::max(mo_lo[i], mo_hi[i]);
        for (int j = 0; j < fs; ++j) {
            int e0 = j - padL, e1 = 1 + j - padL;
            c0.terms.push_back({mod2(e0), floordiv2(e0), j, 0});
            c1.terms.push_back({mod2(e1), floordiv2(e1), j, 0});
        }
        sort_terms(&c0); sort_terms(&c1);
        op.classes = {c0, c1};
    }
    // ---- bottleneck -------------------------------------------------------------------------------
    {
        ConvOp& op = P->bottleneck;
        op.name = "bottleneck";
        op.k = fs; op.cin_tot = F0 * L; op.cout = F0 * (L + 1);
        add_conv_params(fs, op.cin_tot, op.cout, &op.w_param, &op.b_param);
        const int x = t_dec[L - 1];
        const int padL = ctx ? 0 : (fs - 1) / 2;
        op.planes.push_back(direct_view(x, 0, 2, 0, (T[L] + 1) / 2, op.cin_tot));
        op.planes.push_back(direct_view(x, 1, 2, 0, T[L] / 2, op.cin_tot));
        op.plane_grad_tensor = {P->grad_twin[x], P->grad_twin[x]};
        op.plane_slope = {1, 1};
        for (int q = 0; q < 2; ++q) {
            ClassSpec c;
            const int n = (q == 0) ? (Tb + 1) / 2 : Tb / 2;
            c.out = direct_view(t_z, q, 2, 0, n, op.cout);
            c.m_lo = 0; c.m_hi = n;
            for (int j = 0; j < fs; ++j) {
                int e = q + j - padL;
                c.terms.push_back({mod2(e), floordiv2(e), j, 0});
            }
            sort_terms(&c);
            op.classes.push_back(c);
        }
    }
    // ---- up path ----------------------------------------------------------------------------------
    P->up.resize(L);
    P->ups.resize(L);
    int max_ue = 0, max_mid = 0;
    for (int i = 0; i < L; ++i) {
        const int j = L - 1 - i;
        const int src = (i == 0) ? t_z : t_up[i - 1];
        const int Cu = F0 * (L + 1 - i), Cs = F0 * (L - i);
        UpsampleSpec& us = P->ups[i];
        us.src_tensor = src; us.N = Nup[i]; us.nmid = nmid[i]; us.C = Cu;
        us.mid_mode = ctx ? MID_VALID : (learned ? MID_ZERO : MID_CLAMP);
        us.interp_param = -1; us.wsig_tensor = -2;
        if (learned) {                                           // InterpolationLayer.py:19
            char nm[64];
            snprintf(nm, sizeof(nm), "separator/interp_%d", i);
            us.interp_param = add_param(nm, 1, Cu, 1, 1);
            us.wsig_tensor = add_tensor(P, "wsig" + std::to_string(i), 1, Cu, false, false);
        }
        ConvOp& op = P->up[i];
        op.name = "up" + std::to_string(i);
        op.k = mfs; op.cin_tot = Cs + Cu; op.cout = F0 * (L - i);
        add_conv_params(mfs, op.cin_tot, op.cout, &op.w_param, &op.b_param);
        const int padL = ctx ? 0 : (mfs - 1) / 2;
        op.planes.push_back(direct_view(t_dec[j], 0, 1, 0, Td[j], Cs));                       // skip, even rows
        op.planes.push_back(direct_view(t_odd[j], -mo_lo[j], 1, mo_lo[j], std::max(mo_lo[j], mo_hi[j]), Cs));
        op.planes.push_back(direct_view(src, 0, 1, 0, Nup[i], Cu));                           // copied rows
        ViewSpec mid = direct_view(src, 0, 1, 0, nmid[i], Cu);                                // interpolated rows
        mid.kind = PLANE_MID; mid.mid_mode = us.mid_mode; mid.xrows = Nup[i]; mid.blend_tensor = us.wsig_tensor;
        op.planes.push_back(mid);
        max_ue = std::max(max_ue, Nup[i] * Cu);
        max_mid = std::max(max_mid, nmid[i] * Cu);
        op.plane_grad_tensor = {P->grad_twin[t_dec[j]], P->grad_twin[t_odd[j]], -3, -4};   // -3: g_ue, -4: g_mid
        op.plane_slope = {1, 1, 0, 0};
        const int V = Nup[i + 1];
        for (int q = 0; q < 2; ++q) {
            ClassSpec c;
            const int n = (q == 0) ? (V + 1) / 2 : V / 2;
            c.out = direct_view(t_up[i], q, 2, 0, n, op.cout);
            c.m_lo = 0; c.m_hi = n;
            for (int jj = 0; jj < mfs; ++jj) {
                int es = q + jj - padL + cs[i];                   // skip: absolute position parity
                c.terms.push_back({mod2(es), floordiv2(es), jj, 0});
                int eu = q + jj - padL;                           // upsampled: even = copy, odd = interpolated
                c.terms.push_back({2 + mod2(eu), floordiv2(eu), jj, Cs});
            }
            sort_terms(&c);
            op.classes.push_back(c);
        }
    }
    // ---- output convs (OutputLayer.py:8 / :15) ---------------------------------------------------
    for (int s = 0; s < P->nconv; ++s) {
        int w, b;
        add_conv_params(ofs, C + F0, C, &w, &b);
        P->out_w_param.push_back(w);
        P->out_b_param.push_back(b);
    }
    // ---- scratch tensors ----------------------------------------------------------------------------
    P->t_gue = add_tensor(P, "g_ue", max_ue, 1, true, true);
    P->t_gmid = add_tensor(P, "g_mid", std::max(1, max_mid), 1, true, true);
    P->t_dpre_out = add_tensor(P, "dpre_out", P->T_out, P->nconv * C, true, true);
    P->t_outbuf = add_tensor(P, "outbuf", P->T_out * cfg.num_sources, C, true, false);
    for (auto& op : P->up)
        for (auto& g : op.plane_grad_tensor) {
            if (g == -3) g = P->t_gue;
            if (g == -4) g = P->t_gmid;
        }

    // ---- work --------------------------------------------------------------------------------------
    double f = 0;
    for (const auto& op : P->down) f += op_flops(op);
    f += op_flops(P->bottleneck);
    for (const auto& op : P->up) f += op_flops(op);
    f += 2.0 * (double)P->T_out * ofs * (C + F0) * P->nconv * C;
    P->fwd_flops_per_item = f;
    P->dgrad0_flops_per_item = op_flops(P->down[0]);
    return WUN_OK;
}

std::string Plan::describe() const {
    char buf[512];
    std::string s;
    snprintf(buf, sizeof(buf), "Wave-U-Net plan: L=%d F0=%d fs=%d mfs=%d ofs=%d C=%d K=%d context=%d upsampling=%s "
             "output=%s/%s\n  T_in=%lld T_out=%lld feat_rows=%lld params=%lld tensors=%zu\n",
             cfg.num_layers, cfg.num_initial_filters, cfg.filter_size, cfg.merge_filter_size, cfg.output_filter_size,
             cfg.num_channels, cfg.num_sources, cfg.context, cfg.upsampling ? "learned" : "linear",
             cfg.output_type ? "difference" : "direct", cfg.output_activation ? "linear" : "tanh",
             (long long)T_in, (long long)T_out, (long long)Tf, (long long)param_numel, tensors.size());
    s += buf;
    auto dump = [&](const ConvOp& op) {
        snprintf(buf, sizeof(buf), "  %-10s k=%d cin=%d cout=%d  ", op.name.c_str(), op.k, op.cin_tot, op.cout);
        s += buf;
        for (size_t q = 0; q < op.classes.size(); ++q) {
            const auto& c = op.classes[q];
            snprintf(buf, sizeof(buf), "class%zu rows[%d,%d) -> %s  ", q, c.m_lo, c.m_hi,
                     tensors[c.out.tensor].name.c_str());
            s += buf;
        }
        snprintf(buf, sizeof(buf), "%.3f GFLOP/item\n", op_flops(op) * 1e-9);
        s += buf;
    };
    for (const auto& op : down) dump(op);
    dump(bottleneck);
    for (const auto& op : up) dump(op);
    return s;
}

}  // namespace wun
