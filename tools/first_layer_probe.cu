// first_layer_probe.cu - ROUND-2 EXPERIMENT (written at the end of round 1 with no GPU time left: compiles, NOT yet run).
//
// The first down block (C_in = 1 | 2 waveform channels, k = 15, 24 filters; UnetAudioSeparator.py:97-100 with i = 0) is
// memory-shaped work that stays on CUDA cores, but the generic kernels the engine uses for it cost 6 % of the M4 step
// (forward 207 us, weight gradient 457 us in 4 launches) against ~25 us / ~65 us floors.  This probe holds two dedicated
// kernels and runs them next to the engine's on the M4 batch-16 first layer, with a CPU check:
//   first_fwd_kernel   thread = 2 output rows x all N filters in registers; x staged de-interleaved by parity so that the
//                      stride-2 row reads are conflict-free; weights read as warp-broadcast float4 from shared memory
//   first_wgrad_kernel 16 thread groups take the rows of a chunk round-robin; thread = 6 (tap, channel) x 8 filter
//                      register tile (5 shared-memory loads per 48 FMAs); groups reduce through shared memory, one atomic per
//                      weight per CTA; the bias gradient (column sum of g) rides along
//   tools/first_layer_probe [small|m4]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../wave-u-net_b200/csrc/kernels_simt.cu"
#include "../wave-u-net_b200/csrc/kernels_first.cu"

using namespace wun;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)


// ------------------------------------------------------------------------------------------------
struct Problem { int B, T, C, N, k, cs, U; };

static void run(const char* name, Problem p, int iters) {
    const int To = p.T - p.k + 1, Td = (To + 1) / 2;
    const int mo_lo = p.cs / 2, mo_hi = (p.cs + p.U) / 2, n_odd = mo_hi - mo_lo;
    std::vector<float> x((size_t)p.B * p.T * p.C), w((size_t)p.k * p.C * p.N), bias(p.N);
    std::vector<float> gdec((size_t)p.B * Td * p.N), godd((size_t)p.B * (n_odd + 1) * p.N);
    unsigned s = 4242u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : x) v = rnd();
    for (auto& v : w) v = rnd() * 0.3f;
    for (auto& v : bias) v = rnd() * 0.1f;
    for (auto& v : gdec) v = rnd() * 1e-3f;
    for (auto& v : godd) v = rnd() * 1e-3f;
    float *dx, *dw, *db, *ddec, *dodd, *dgd, *dgo, *ddw, *ddb;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4)); CK(cudaMalloc(&db, bias.size() * 4));
    CK(cudaMalloc(&ddec, gdec.size() * 4)); CK(cudaMalloc(&dodd, godd.size() * 4));
    CK(cudaMalloc(&dgd, gdec.size() * 4)); CK(cudaMalloc(&dgo, godd.size() * 4));
    CK(cudaMalloc(&ddw, w.size() * 4)); CK(cudaMalloc(&ddb, p.N * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dgd, gdec.data(), gdec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dgo, godd.data(), godd.size() * 4, cudaMemcpyHostToDevice));

    // ---- the engine's launches for this layer (plan.cpp down0: two parity planes of the waveform, classes dec / odd) ----
    ConvLaunch CL;
    memset(&CL, 0, sizeof(CL));
    CL.nplanes = 2;
    for (int par = 0; par < 2; ++par) {
        PlaneView& P = CL.planes[par];
        P.base = dx + par * p.C; P.bstride = (long long)p.T * p.C; P.rstride = 2 * p.C;
        P.r_lo = 0; P.r_hi = (par == 0) ? (p.T + 1) / 2 : p.T / 2; P.C = p.C; P.kind = PLANE_DIRECT;
    }
    CL.ncls = 2; CL.N = p.N; CL.w_sk = p.N; CL.w_sn = 1; CL.W = dw; CL.bias = db; CL.epilogue = EPI_BIAS_LRELU; CL.batch = p.B;
    int nt = 0;
    for (int q = 0; q < 2; ++q) {
        OutView& O = CL.cls[q];
        O.base = (q == 0) ? ddec : dodd - (long long)mo_lo * p.N;
        O.bstride = (q == 0) ? (long long)Td * p.N : (long long)n_odd * p.N;
        O.rstride = p.N; O.m_lo = (q == 0) ? 0 : mo_lo; O.m_hi = (q == 0) ? Td : mo_hi;
        O.term_begin = nt;
        for (int par = 0; par < 2; ++par)
            for (int j = 0; j < p.k; ++j) { int e = q + j; if ((e & 1) == par) CL.terms[nt++] = {par, e >> 1, j * p.C * p.N}; }
        O.term_end = nt;
        CL.max_rows = max(CL.max_rows, O.m_hi - O.m_lo);
    }
    std::vector<WgradLaunch> WLs;
    for (int q = 0; q < 2; ++q)
        for (int par = 0; par < 2; ++par) {
            WgradLaunch W;
            memset(&W, 0, sizeof(W));
            W.plane = CL.planes[par];
            W.dpre.kind = PLANE_DIRECT; W.dpre.C = p.N; W.dpre.rstride = p.N;
            W.dpre.base = (q == 0) ? dgd : dgo - (long long)mo_lo * p.N;
            W.dpre.bstride = (q == 0) ? (long long)Td * p.N : (long long)n_odd * p.N;
            W.m_lo = (q == 0) ? 0 : mo_lo; W.m_hi = (q == 0) ? Td : mo_hi;
            W.dpre.r_lo = W.m_lo; W.dpre.r_hi = W.m_hi;
            for (int j = 0; j < p.k; ++j) { int e = q + j; if ((e & 1) == par) { W.d[W.nterms] = e >> 1; W.woff[W.nterms] = j * p.C * p.N; ++W.nterms; } }
            W.N = p.N; W.w_sk = p.N; W.w_sn = 1; W.dW = ddw; W.scale = 1.f; W.batch = p.B;
            if (W.m_hi > W.m_lo && W.nterms) WLs.push_back(W);
        }

    FirstLayer FL;
    memset(&FL, 0, sizeof(FL));
    FL.x = dx; FL.x_bstride = (long long)p.T * p.C; FL.T = p.T; FL.k = p.k; FL.pad_left = 0;
    FL.dec = ddec; FL.dec_bstride = (long long)Td * p.N; FL.Td = Td;
    FL.odd = dodd; FL.odd_bstride = (long long)n_odd * p.N; FL.mo_lo = mo_lo; FL.mo_hi = mo_hi;
    FL.W = dw; FL.bias = db; FL.batch = p.B;
    FirstWgrad FW;
    FW.L = FL; FW.L.dec = dgd; FW.L.odd = dgo; FW.dW = ddw; FW.db = ddb; FW.scale = 1.f;
    FW.rows_per_cta = 0;
    if (!first_layer_supported(p.C, p.N, p.k)) { printf("[%s] shape not supported by the dedicated kernels\n", name); return; }
    auto new_fwd = [&]() { launch_first_fwd(FL, p.C, p.N, 0); };
    auto new_wgrad = [&]() { launch_first_wgrad(FW, p.C, p.N, 0); };

    // ---- forward check ----
    auto check_fwd = [&](const char* which) {
        std::vector<float> dec((size_t)p.B * Td * p.N), odd((size_t)p.B * (n_odd + 1) * p.N);
        CK(cudaMemcpy(dec.data(), ddec, dec.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(odd.data(), dodd, odd.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0; long long bad = 0, tested = 0;
        const int stride = To > 4000 ? 53 : 1;
        for (int b = 0; b < p.B; ++b)
            for (int a = 0; a < To; a += stride) {
                const bool even = (a & 1) == 0;
                if (!even && !(a >= p.cs && a < p.cs + p.U)) continue;
                for (int n = 0; n < p.N; ++n) {
                    double acc = bias[n];
                    for (int j = 0; j < p.k; ++j)
                        for (int c = 0; c < p.C; ++c) acc += (double)x[((size_t)b * p.T + a + j) * p.C + c] * w[((size_t)j * p.C + c) * p.N + n];
                    const double ref = acc > 0 ? acc : 0.2 * acc;
                    const float got = even ? dec[((size_t)b * Td + a / 2) * p.N + n] : odd[((size_t)b * n_odd + ((a - 1) / 2 - mo_lo)) * p.N + n];
                    const double err = fabs(got - ref) / (fabs(ref) + 1e-2);
                    if (!(err < 1e-5)) ++bad;
                    worst = fmax(worst, err); ++tested;
                }
            }
        printf("[%s] fwd   %-7s %s tested=%lld bad=%lld worst=%.2e\n", name, which, bad ? "FAIL" : "PASS", tested, bad, worst);
    };
    auto check_wgrad = [&](const char* which, bool with_bias) {
        std::vector<float> gw(w.size()), gb(p.N);
        CK(cudaMemcpy(gw.data(), ddw, gw.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gb.data(), ddb, p.N * 4, cudaMemcpyDeviceToHost));
        std::vector<double> ref(w.size(), 0.0), refb(p.N, 0.0);
        for (int b = 0; b < p.B; ++b)
            for (int a = 0; a < To; ++a) {
                const bool even = (a & 1) == 0;
                if (!even && !(a >= p.cs && a < p.cs + p.U)) continue;
                const float* g = even ? &gdec[((size_t)b * Td + a / 2) * p.N] : &godd[((size_t)b * n_odd + ((a - 1) / 2 - mo_lo)) * p.N];
                for (int n = 0; n < p.N; ++n) refb[n] += g[n];
                for (int j = 0; j < p.k; ++j)
                    for (int c = 0; c < p.C; ++c) {
                        const double xv = x[((size_t)b * p.T + a + j) * p.C + c];
                        for (int n = 0; n < p.N; ++n) ref[((size_t)j * p.C + c) * p.N + n] += xv * g[n];
                    }
            }
        double num = 0, den = 0, numb = 0, denb = 0;
        for (size_t i = 0; i < ref.size(); ++i) { num += (gw[i] - ref[i]) * (gw[i] - ref[i]); den += ref[i] * ref[i]; }
        for (int n = 0; n < p.N; ++n) { numb += (gb[n] - refb[n]) * (gb[n] - refb[n]); denb += refb[n] * refb[n]; }
        const double rel = sqrt(num / den), relb = sqrt(numb / fmax(denb, 1e-300));
        printf("[%s] wgrad %-7s %s rel_l2=%.2e%s\n", name, which, (rel < 1e-4 && (!with_bias || relb < 1e-4)) ? "PASS" : "FAIL", rel,
               with_bias ? (relb < 1e-4 ? "  (bias grad ok)" : "  (bias grad WRONG)") : "");
    };
    auto time_it = [&](const char* what, auto fn, double flops) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        fn();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) fn();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
        printf("[%s] %-22s %8.1f us  %.1f TFLOP/s\n", name, what, ms * 1e3, flops / (ms * 1e-3) * 1e-12);
    };
    const double flops = 2.0 * p.B * ((double)Td + n_odd) * p.k * p.C * p.N;
    const bool do_check = (long long)p.B * p.T < 4000000;

    CK(cudaMemset(ddec, 0xFF, gdec.size() * 4)); CK(cudaMemset(dodd, 0xFF, godd.size() * 4));
    launch_plane_conv_simt(CL, 0);
    CK(cudaDeviceSynchronize());
    if (do_check) check_fwd("engine");
    CK(cudaMemset(ddec, 0xFF, gdec.size() * 4)); CK(cudaMemset(dodd, 0xFF, godd.size() * 4));
    new_fwd();
    { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("[%s] new fwd ERROR %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (do_check) check_fwd("new");

    CK(cudaMemset(ddw, 0, w.size() * 4)); CK(cudaMemset(ddb, 0, p.N * 4));
    for (auto& W : WLs) launch_plane_wgrad_simt(W, 0);
    CK(cudaDeviceSynchronize());
    if (do_check) check_wgrad("engine", false);
    CK(cudaMemset(ddw, 0, w.size() * 4)); CK(cudaMemset(ddb, 0, p.N * 4));
    new_wgrad();
    { cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("[%s] new wgrad ERROR %s\n", name, cudaGetErrorString(e)); exit(3); } }
    if (do_check) check_wgrad("new", true);

    if (iters > 0) {
        time_it("fwd   engine (generic)", [&]() { launch_plane_conv_simt(CL, 0); }, flops);
        time_it("fwd   new", new_fwd, flops);
        time_it("wgrad engine (4 launches)", [&]() { for (auto& W : WLs) launch_plane_wgrad_simt(W, 0); }, flops);
        time_it("wgrad new (+bias grad)", new_wgrad, flops);
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(ddec); cudaFree(dodd); cudaFree(dgd); cudaFree(dgo); cudaFree(ddw); cudaFree(ddb);
}

int main(int argc, char** argv) {
    const char* which = argc > 1 ? argv[1] : "all";
    auto want = [&](const char* n) { return !strcmp(which, "all") || !strcmp(which, n); };
    if (want("small")) {
        run("stereo_small", {2, 3000, 2, 24, 15, 700, 901}, 0);
        run("mono_small",   {2, 2111, 1, 24, 15, 301, 600}, 0);
        run("k9_small",     {1, 1500, 2, 24, 9, 100, 333}, 0);
    }
    if (want("m4")) run("m4_down0_b16", {16, 147443, 2, 24, 15, 65518, 16393}, 10);   // down0: odd rows [32759, 40955)
    return 0;
}
