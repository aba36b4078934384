// umma_probe.cu - standalone check + timing of the tcgen05 plane-conv kernel (wave-u-net_b200/csrc/kernels_umma.cu)
// on "down block" shaped problems (conv k taps, valid, decimated outputs + odd outputs in a window).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_probe tools/umma_probe.cu
//   ./tools/umma_probe            (run under gpurun; prints PASS/FAIL per case and TFLOP/s per timed layer)
// The CPU reference uses the same hi/lo bf16 split of both operands (3 of the 4 partial products), in double,
// so any mismatch beyond fp32 accumulation noise is a kernel / descriptor bug.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../wave-u-net_b200/csrc/kernels_umma.cu"

using namespace wun;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Problem {
    int B, T, Cin, Cout, fs, MT, nsplit;
    int cs, U;      // skip window [cs, cs+U) in full-rate output coordinates
};

static double run_case(const char* name, Problem p, bool check, int timing_iters) {
    const int To = p.T - p.fs + 1, Td = (To + 1) / 2;
    const int mo_lo = p.cs / 2, mo_hi = (p.cs + p.U) / 2, n_odd = mo_hi - mo_lo;
    std::vector<float> x((size_t)p.B * p.T * p.Cin), w((size_t)p.fs * p.Cin * p.Cout), bias(p.Cout);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : x) v = rnd();
    for (auto& v : w) v = rnd() * 0.2f;
    for (auto& v : bias) v = rnd() * 0.1f;
    float *dx, *dw, *db, *ddec, *dodd;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4)); CK(cudaMalloc(&db, bias.size() * 4));
    CK(cudaMalloc(&ddec, (size_t)p.B * Td * p.Cout * 4)); CK(cudaMalloc(&dodd, (size_t)p.B * (n_odd + 1) * p.Cout * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(ddec, 0xFF, (size_t)p.B * Td * p.Cout * 4));
    CK(cudaMemset(dodd, 0xFF, (size_t)p.B * (n_odd + 1) * p.Cout * 4));

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
    if (p.MT > 0 && p.MT != ch.MT) {       // override the tiling for the sweep
        int span = ch.rows_alloc - ch.MT * 128;
        ch.MT = p.MT; ch.rows_alloc = p.MT * 128 + span;
        int tm = 32; while (tm < ch.MT * ch.NPAD) tm *= 2;
        ch.tmem_cols = tm;
        if (ch.persistent) {
            if (2 * ch.MT * ch.NPAD > 512) ch.persistent = 0;
            else { tm = 32; while (tm < 2 * ch.MT * ch.NPAD) tm *= 2; ch.tmem_cols = tm; }
        }
    }
    printf("[%s] persistent=%d TB=%d nbs=%d\n", name, ch.persistent, ch.TB, ch.nbs);
    const int NPAD = ch.NPAD;
    uint8_t* arena;
    CK(cudaMalloc(&arena, ch.pack_bytes));
    UmmaLaunch L;
    UmmaPackLaunch PL;
    CK(umma_build(CL, ch, arena, &L, &PL));
    CK(launch_umma_pack(PL, 0));
    std::vector<uint8_t*> packs;
    packs.push_back(arena);
    CK(cudaDeviceSynchronize());
    printf("[%s] B=%d T=%d Cin=%d Cout=%d fs=%d MT=%d nsplit=%d NPAD=%d rows_alloc=%d smem=%zu tmem=%d\n", name, p.B, p.T,
           p.Cin, p.Cout, p.fs, L.MT, L.nsplit, NPAD, L.rows_alloc, umma_smem_bytes(L), L.tmem_cols);
    CK(launch_plane_conv_umma(L, 0));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); }

    double worst = 0;
    if (check) {
        std::vector<float> dec((size_t)p.B * Td * p.Cout), odd((size_t)p.B * (n_odd + 1) * p.Cout);
        CK(cudaMemcpy(dec.data(), ddec, dec.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(odd.data(), dodd, odd.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<float> xh(x.size()), xl(x.size()), wh(w.size()), wl(w.size());
        for (size_t i = 0; i < x.size(); ++i) { xh[i] = bf16_round(x[i]); xl[i] = bf16_round(x[i] - xh[i]); }
        for (size_t i = 0; i < w.size(); ++i) { wh[i] = bf16_round(w[i]); wl[i] = bf16_round(w[i] - wh[i]); }
        long long nbad = 0, ntested = 0;
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
                    if (!(err < 2e-4)) { if (nbad < 5) printf("   mismatch b=%d a=%d n=%d got=%g ref=%g\n", b, a, n, got, ref); ++nbad; }
                    if (err > worst || err != err) worst = err;
                    ++ntested;
                }
            }
        printf("[%s] %s  tested=%lld bad=%lld worst_rel=%.3e\n", name, nbad == 0 ? "PASS" : "FAIL", ntested, nbad, worst);
    }
    double tf = 0;
    if (timing_iters > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) CK(launch_plane_conv_umma(L, 0));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < timing_iters; ++i) CK(launch_plane_conv_umma(L, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= timing_iters;
        const double flops = 2.0 * p.B * ((double)Td + n_odd) * p.fs * p.Cin * p.Cout;
        tf = flops / (ms * 1e-3) * 1e-12;
        printf("[%s] time %.1f us  %.1f useful TFLOP/s (x3 MMAs issued, padded K/N not counted)\n", name, ms * 1e3, tf);
#ifdef WUN_UMMA_TIMING
        unsigned long long tt[16];
        CK(cudaMemcpyFromSymbol(tt, g_umma_timing, sizeof(tt)));
        const double n = (double)tt[0];
        printf("   per-CTA cycles: total %.0f prologue %.0f | mma loop %.0f (wait slab %.0f, wait B %.0f) | conv wait-empty %.0f fill %.0f | epi wait-acc %.0f epi %.0f | loader wait %.0f  (CTAs %.0f)\n",
               tt[1] / n, tt[2] / n, tt[3] / n, tt[4] / n, tt[5] / n, tt[6] / n, tt[7] / n, tt[8] / n, tt[9] / n, tt[10] / n, n / (timing_iters + 4));
        memset(tt, 0, sizeof(tt));
        CK(cudaMemcpyToSymbol(g_umma_timing, tt, sizeof(tt)));
#endif
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(ddec); cudaFree(dodd);
    for (auto q : packs) cudaFree(q);
    return worst;
}

// ------------------------------------------------------------------------------------------------
// wgrad check / timing on the same "down block" geometry
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
    std::vector<UmmaWgradLaunch> launches(1, WL);
    const WgGroup& W0 = WL.grp[0];
    printf("[%s] wgrad B=%d T=%d Cin=%d Cout=%d swap=%d NT=%d mtiles=%d ntiles=%d taps/cta=%d tapsets=%d chunks/cta=%d grid=(%d,%d,%d)\n", name, p.B, p.T,
           p.Cin, p.Cout, W0.swap, W0.NT, W0.n_mtiles, W0.n_ntiles, W0.taps_per_cta, W0.n_tapsets, W0.chunks_per_cta, WL.grid_x, WL.grid_y, WL.grid_z);
    CK(cudaMemset(ddw, 0, wn * 4));
    for (auto& W : launches) CK(launch_wgrad_umma(W, 0));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] WGRAD KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); }
    if (check) {
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
        double num = 0, den = 0, worst = 0;
        for (size_t i = 0; i < wn; ++i) { num += (dw[i] - ref[i]) * (dw[i] - ref[i]); den += ref[i] * ref[i]; worst = fmax(worst, fabs(dw[i] - ref[i])); }
        const double rel = sqrt(num / den);
        printf("[%s] wgrad %s rel_l2=%.3e max_abs_err=%.3e (|ref| rms %.3e)\n", name, rel < 1e-4 ? "PASS" : "FAIL", rel, worst, sqrt(den / wn));
    }
    if (timing_iters > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (auto& W : launches) CK(launch_wgrad_umma(W, 0));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < timing_iters; ++i) for (auto& W : launches) CK(launch_wgrad_umma(W, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= timing_iters;
        const double flops = 2.0 * p.B * ((double)Td + n_odd) * p.fs * p.Cin * p.Cout;
        printf("[%s] wgrad time %.1f us (1 launch)  %.1f useful TFLOP/s\n", name, ms * 1e3, flops / (ms * 1e-3) * 1e-12);
#ifdef WUN_UMMA_TIMING
        unsigned long long tt[16];
        CK(cudaMemcpyFromSymbol(tt, g_umma_timing, sizeof(tt)));
        const double n = (double)tt[0];
        printf("   per-CTA cycles: mma loop %.0f (wait full %.0f) | conv wait-empty %.0f | epi wait-acc %.0f epi %.0f | chunks/CTA %.1f stages %d (CTAs %.0f)\n",
               tt[3] / n, tt[4] / n, tt[6] / n, tt[8] / n, tt[9] / n, tt[7] / n, WL.nstages, n / (timing_iters + 1));
        memset(tt, 0, sizeof(tt));
        CK(cudaMemcpyToSymbol(g_umma_timing, tt, sizeof(tt)));
#endif
    }
    cudaFree(dx); cudaFree(dgd); cudaFree(dgo); cudaFree(ddw);
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    if (argc > 2 && !strcmp(argv[1], "only")) {      // single timed layer, for ncu
        run_case("down3_only", {16, 18419, 72, 96, 15, atoi(argv[2]), 1, 8174, 2057}, false, 3);
        return 0;
    }
    //                      B   T     Cin Cout fs MT ns   cs   U
    run_case("tiny",       { 1, 300,   16, 16,  3, 1, 1,   40, 101}, true, 0);
    run_case("taps15",     { 2, 1500,  32, 48, 15, 1, 1,  200, 401}, true, 0);
    run_case("cin24_mt2",  { 2, 1500,  24, 48, 15, 2, 1,  200, 401}, true, 0);
    run_case("c72_n96",    { 2, 3000,  72, 96, 15, 0, 1,  500, 801}, true, 0);
    run_case("n72pad_mt2", { 2, 1000,  48, 72, 15, 2, 1,  100, 301}, true, 0);
    run_case("nsplit",     { 2, 300,   64, 288, 15, 1, 2,  50, 101}, true, 0);
    run_wgrad("wg_tiny",   { 1, 300,   16, 16,  3, 1, 1,   40, 101}, true, 0);
    run_wgrad("wg_taps15", { 2, 1500,  32, 48, 15, 1, 1,  200, 401}, true, 0);
    run_wgrad("wg_c72n96", { 2, 3000,  72, 96, 15, 1, 1,  500, 801}, true, 0);
    run_wgrad("wg_c24",    { 2, 1500,  24, 48, 15, 1, 1,  200, 401}, true, 0);
    run_wgrad("wg_wide",   { 2, 300,  264, 288, 15, 1, 1,  50, 101}, true, 0);
    if (argc > 1 && !strcmp(argv[1], "notime")) return 0;
    run_wgrad("wg_down1",  {16, 73715, 24, 48, 15, 1, 1, 32750, 8201}, false, 10);
    run_wgrad("wg_down2",  {16, 36851, 48, 72, 15, 1, 1, 16366, 4105}, false, 10);
    run_wgrad("wg_down3",  {16, 18419, 72, 96, 15, 1, 1, 8174, 2057}, false, 10);
    run_wgrad("wg_down5",  {16, 4595, 120, 144, 15, 1, 1, 2030, 521}, false, 10);
    run_wgrad("wg_down8",  {16, 563, 192, 216, 15, 1, 1, 242, 69}, false, 10);
    if (argc > 2 && !strcmp(argv[1], "only")) {      // single timed layer, for ncu
        int mt = atoi(argv[2]);
        run_case("down3_only", {16, 18419, 72, 96, 15, mt, 1, 8174, 2057}, false, 3);
        return 0;
    }
    // M4 layers at B=16 (T = input rows of the layer)
    if (argc > 1 && !strcmp(argv[1], "dgrad")) {       // shapes of the down-block dgrads (few output channels, many rows)
        run_case("dg1_like", {16, 147443, 48, 24, 15, 0, 1, 65500, 16401}, false, 10);
        run_case("dg2_like", {16, 73715, 72, 48, 15, 0, 1, 32750, 8201}, false, 10);
        run_case("dg3_like", {16, 36851, 96, 72, 15, 0, 1, 16366, 4105}, false, 10);
        return 0;
    }
    run_case("down1",      {16, 73715, 24, 48, 15, 0, 1, 32750, 8201}, true, 20);
    run_case("down2_mt2",  {16, 36851, 48, 72, 15, 0, 1, 16366, 4105}, false, 20);
    run_case("down3_mt2",  {16, 18419, 72, 96, 15, 0, 1, 8174, 2057}, false, 20);
    run_case("down4_mt2",  {16, 9203,  96, 120, 15, 0, 1, 4078, 1033}, false, 20);
    run_case("down7_mt2",  {16, 1139, 168, 192, 15, 0, 1, 494, 137}, false, 20);
    return 0;
}
