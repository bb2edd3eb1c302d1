// Scratch harness (not part of the product): times yh_v2_train / yh_v2_postprocess through the C ABI
// under the bench's regime (CUDA graph of 8 launches over 8 rotating buffer sets), with the
// kernels compiled into this binary so -DYH_X_* experiment switches can be flipped per build.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I <csrc> -o tb scratch/trainbench.cu <csrc>/yh_api.cu <csrc>/yh_train.cu <csrc>/yh_nms.cu  (see scratch/mk.sh)
#include <cuda_runtime.h>
#include <stdio.h>
#include "yolohead.h"

#include <stdlib.h>
#include <vector>
#include <random>
#include <functional>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

#ifdef YH_X_TRACE
extern "C" int yh_x_trace_copy(unsigned long long*, int);
extern "C" int yh_x_ntrace_copy(unsigned long long*, int);
extern "C" int yh_x_rcycles_copy(unsigned int*, int);
extern "C" int yh_x_tiletrace_copy(unsigned long long*, int);
#endif

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256, S = 13, A = 5, C = 20, R = 8;
    const float mu = argc > 2 ? atof(argv[2]) : -1.563f;  // objectness shift: ~50 of 845 pass conf 0.5
    const size_t floats = (size_t)N * S * S * A * (5 + C);
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> hy(floats);
    for (size_t i = 0; i < floats; ++i) hy[i] = nd(rng) + ((i % 25) == 4 ? mu : 0.f);
    std::vector<YhGt> gt;
    std::vector<int> off(N + 1, 0);
    for (int n = 0; n < N; ++n) {
        const int kmin = argc > 3 ? atoi(argv[3]) : 1, kmax = argc > 4 ? atoi(argv[4]) : 5;  // boxes per image
        const int k = kmin + rng() % (kmax - kmin + 1);
        for (int j = 0; j < k; ++j) {
            YhGt r;
            const float w = 416.f * (0.05f + 0.55f * (rng() % 1000) / 1000.f), h = 416.f * (0.05f + 0.55f * (rng() % 1000) / 1000.f);
            const float cx = w / 2 + (416.f - w) * (rng() % 1000) / 1000.f, cy = h / 2 + (416.f - h) * (rng() % 1000) / 1000.f;
            r.img = n; r.cx = (int)(cx / 32); r.cy = (int)(cy / 32); r.cls = rng() % C;
            r.stx = cx / 32 - r.cx; r.sty = cy / 32 - r.cy; r.tw = w / 32; r.th = h / 32;
            r.x1 = cx - w / 2; r.y1 = cy - h / 2; r.x2 = cx + w / 2; r.y2 = cy + h / 2;
            gt.push_back(r);
        }
        off[n + 1] = (int)gt.size();
    }
    const int M = (int)gt.size();
    const float anchors[10] = {1.3221f, 1.73145f, 3.19275f, 4.00944f, 5.05587f, 8.09892f, 9.47112f, 4.84053f, 11.2364f, 10.0071f};
    const float lam[5] = {5, 5, 1, .5f, 1};
    const int MAXO = 128;
    int post_flags = 0;
    bool train_ov = false;

    struct Set { float *y, *dy, *terms, *loss, *obox, *oconf, *oscore; YhGt* gt; int *off, *kidx, *kcnt, *olab; void *ws, *pws; };
    std::vector<Set> sets(R);
    const size_t pws_bytes = yh_postprocess_workspace_bytes(N, S * S * A);
    for (auto& s : sets) {
        CK(cudaMalloc(&s.y, floats * 4)); CK(cudaMalloc(&s.dy, floats * 4));
        CK(cudaMalloc(&s.gt, M * sizeof(YhGt))); CK(cudaMalloc(&s.off, (N + 1) * 4));
        CK(cudaMalloc(&s.terms, 32)); CK(cudaMalloc(&s.loss, 4));
        CK(cudaMalloc(&s.ws, yh_train_workspace_bytes())); CK(cudaMemset(s.ws, 0, yh_train_workspace_bytes()));
        CK(cudaMalloc(&s.pws, pws_bytes));
        CK(cudaMalloc(&s.kidx, (size_t)N * MAXO * 4)); CK(cudaMalloc(&s.kcnt, N * 4)); CK(cudaMalloc(&s.olab, (size_t)N * MAXO * 4));
        CK(cudaMalloc(&s.obox, (size_t)N * MAXO * 16)); CK(cudaMalloc(&s.oconf, (size_t)N * MAXO * 4)); CK(cudaMalloc(&s.oscore, (size_t)N * MAXO * 4));
        CK(cudaMemcpy(s.y, hy.data(), floats * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s.gt, gt.data(), M * sizeof(YhGt), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s.off, off.data(), (N + 1) * 4, cudaMemcpyHostToDevice));
    }
    cudaStream_t st; CK(cudaStreamCreate(&st));

    auto run = [&](const char* name, double bytes, std::function<int(Set&)> launch) {
        for (auto& s : sets) { int rc = launch(s); if (rc) { printf("%s failed: %s\n", name, yh_last_error()); exit(1); } }
        CK(cudaStreamSynchronize(st));
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        for (auto& s : sets) launch(s);
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ge, st));
        CK(cudaStreamSynchronize(st));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int reps = 100;
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / (reps * R);
        printf("%-36s %7.2f us  %7.0f GB/s (%.1f%% of 6552.6)\n", name, us, bytes / us / 1e3, bytes / us / 1e3 / 65.526);
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    };
    auto train = [&](Set& s) {
        return (train_ov ? yh_v2_train_overlapped : yh_v2_train)(s.y, N, S, S, A, C, anchors, 416.f, 416.f, s.gt, s.off, M, M, lam, s.dy, s.terms,
                                                                 s.loss, nullptr, nullptr, s.ws, yh_train_workspace_bytes(), st);
    };
    auto post = [&](Set& s) {
        return yh_v2_postprocess(s.y, N, S, S, A, C, anchors, 416.f, 416.f, 0.5f, 0.45f, post_flags, MAXO, s.kidx, s.kcnt, s.obox, s.oconf,
                                 nullptr, s.olab, s.oscore, s.pws, pws_bytes, st);
    };
    const double tb = 2.0 * floats * 4 + 48.0 * M;
    std::vector<void*> fws(R);
    const size_t fws_bytes = yh_train_post_workspace_bytes(N, S, S, A, C);
    for (int i = 0; i < R; ++i) { CK(cudaMalloc(&fws[i], fws_bytes)); CK(cudaMemset(fws[i], 0, fws_bytes)); }
    int fused_flags = 0;
    auto fused = [&](Set& s) {
        return yh_v2_train_post(s.y, N, S, S, A, C, anchors, 416.f, 416.f, s.gt, s.off, M, M, lam, s.dy, s.terms, s.loss, nullptr, nullptr,
                                0.5f, 0.45f, fused_flags, MAXO, s.kidx, s.kcnt, s.obox, s.oconf, nullptr, s.olab, s.oscore, nullptr,
                                fws[&s - &sets[0]], fws_bytes, st);
    };
    run("fused (stream-ordered)", tb + floats * 4, fused);
    fused_flags = YH_STEP_OVERLAPPED;
    run("fused (overlapped chain)", tb + floats * 4, fused);
    fused_flags = 0;
#ifdef YH_X_TRACE
    {
        for (int rep = 0; rep < 3; ++rep) { fused(sets[rep]); CK(cudaStreamSynchronize(st)); }
        std::vector<unsigned long long> tr(4096 * 16);
        yh_x_ntrace_copy(tr.data(), 4096 * 16);
        const int G = N;
        unsigned long long t0 = ~0ull; for (int b = 0; b < G; ++b) t0 = std::min(t0, tr[b * 16]);
        const char* nm[16] = {"fused start", "copies issued", "lists arrived", "-", "B rank+decode done", "-", "-", "D pairs done", "D greedy done", "records start", "records done", "sums published", "-", "E emit done", "-", "-"};
        for (int sl = 0; sl < 14; ++sl) {
            if (nm[sl][0] == '-') continue;
            std::vector<double> v; for (int b = 0; b < G; ++b) if (tr[b * 16 + sl] >= t0) v.push_back((double)(tr[b * 16 + sl] - t0));
            if (v.empty()) continue;
            std::sort(v.begin(), v.end());
            printf("  fused %-20s n=%3zu min %6.0f  p10 %6.0f  med %6.0f  p90 %6.0f  max %6.0f ns\n", nm[sl], v.size(), v[0], v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back());
        }
    }
#endif
    run("train", tb, train);
    run("post", 1.0 * floats * 4, post);
    run("train+post", tb + floats * 4, [&](Set& s) { int rc = train(s); return rc ? rc : post(s); });
    post_flags = YH_POST_INPUT_READY;
    run("train+post (input ready)", tb + floats * 4, [&](Set& s) { int rc = train(s); return rc ? rc : post(s); });
    train_ov = true;
    run("train+post (both overlapped)", tb + floats * 4, [&](Set& s) { int rc = train(s); return rc ? rc : post(s); });
    run("post+train (both overlapped)", tb + floats * 4, [&](Set& s) { int rc = post(s); return rc ? rc : train(s); });
    run("train (overlapped, back to back)", tb, train);
    post_flags = 0;
    run("train overlapped + post plain", tb + floats * 4, [&](Set& s) { int rc = train(s); return rc ? rc : post(s); });
    post_flags = YH_POST_INPUT_READY;
    train_ov = false;
    post_flags = 0;
#ifdef YH_X_TRACE
    {
        CK(cudaMemset(sets[0].dy, 0, 16));
        for (int rep = 0; rep < 3; ++rep) { train(sets[rep]); CK(cudaStreamSynchronize(st)); }
        {   // per-tile timeline of the CTAs that share SM 0..1 (multi-tile launches): start | dense pass done | records done
            std::vector<unsigned long long> tt(4096 * 32);
            yh_x_tiletrace_copy(tt.data(), 4096 * 32);
            unsigned long long tmin = ~0ull;
            for (int b = 0; b < 4096; ++b) if (tt[b * 32] && tt[b * 32 + 1]) tmin = std::min(tmin, tt[b * 32 + 1]);
            for (int sm = 0; sm < 2; ++sm)
                for (int b = 0; b < 4096; ++b) {
                    if (tt[b * 32] != (unsigned long long)sm + 1) continue;
                    printf("  sm %d cta %4d:", sm, b);
                    for (int k = 0; k < 10 && tt[b * 32 + 1 + 3 * k]; ++k)
                        printf("  [%5.1f %5.1f %5.1f]", (tt[b * 32 + 1 + 3 * k] - tmin) / 1e3, (tt[b * 32 + 2 + 3 * k] - tmin) / 1e3, (tt[b * 32 + 3 + 3 * k] - tmin) / 1e3);
                    printf("\n");
                }
        }
        std::vector<unsigned long long> tr(4096 * 24);
        yh_x_trace_copy(tr.data(), 4096 * 24);
        int G = 0; while (G < 4095 && tr[G * 24] != 0) ++G;  // CTAs that left a trace
        unsigned long long t0 = ~0ull; for (int b = 0; b < G; ++b) t0 = std::min(t0, tr[b * 24]);
        const char* nm[24] = {"t0 start", "t0 dense done", "t0 after sync B", "t0 tile loop done", "-", "-", "-", "-",
                              "rw start", "rw before sync B", "rw after sync B", "rw tile loop done", "rw offsets arrived", "rw list built", "rw records done", "-",
                              "pub start", "-", "-", "pub tile loop done", "pub publish begins", "pub fence done", "pub ticket done", "pub final (last only)"};
        for (int sl = 0; sl < 24; ++sl) {
            if (nm[sl][0] == '-') continue;
            std::vector<double> v; for (int b = 0; b < G; ++b) if (tr[b * 24 + sl] >= t0 && tr[b*24+sl] < t0 + 1000000) v.push_back((double)(tr[b * 24 + sl] - t0));
            if (v.empty()) continue; std::sort(v.begin(), v.end());
            printf("  %-22s n=%3zu min %6.0f  p10 %6.0f  med %6.0f  p90 %6.0f  max %6.0f ns\n", nm[sl], v.size(), v[0], v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back());
        }
    }
#endif
#ifdef YH_X_TRACE
    {
        std::vector<unsigned int> rc(4096 * 16);
        yh_x_rcycles_copy(rc.data(), 4096 * 16);
        std::vector<unsigned int> v; for (unsigned x : rc) if (x) v.push_back(x);
        std::sort(v.begin(), v.end());
        if (!v.empty()) printf("  process_record cycles: n=%zu min %u p10 %u med %u p90 %u max %u\n", v.size(), v[0], v[v.size()/10], v[v.size()/2], v[v.size()*9/10], v.back());
    }
    {
        for (int rep = 0; rep < 3; ++rep) { post(sets[rep]); CK(cudaStreamSynchronize(st)); }
        std::vector<unsigned long long> tr(4096 * 16);
        yh_x_ntrace_copy(tr.data(), 4096 * 16);
        const int G = N;
        unsigned long long t0 = ~0ull; for (int b = 0; b < G; ++b) t0 = std::min(t0, tr[b * 16]);
        const char* nm[16] = {"nms start", "init done", "A issued+arrived", "rank done", "B decode done", "D1 done", "D mask zeroed", "D pairs done", "D fixed point done", "-", "-", "-", "D done", "E emit done", "-", "-"};
        for (int sl = 0; sl < 14; ++sl) {
            if (nm[sl][0] == '-') continue;
            std::vector<double> v; for (int b = 0; b < G; ++b) v.push_back((double)(tr[b * 16 + sl] - t0));
            std::sort(v.begin(), v.end());
            printf("  %-22s n=%3zu min %6.0f  p10 %6.0f  med %6.0f  p90 %6.0f  max %6.0f ns\n", nm[sl], v.size(), v[0], v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back());
        }
    }
#endif
#ifdef YH_X_TRACE
    {   // the chain of the timed region: train head, then post-process with YH_POST_INPUT_READY, on one time base
        post_flags = YH_POST_INPUT_READY;
        for (int rep = 0; rep < 3; ++rep) { train(sets[rep]); post(sets[rep]); CK(cudaStreamSynchronize(st)); }
        post_flags = 0;
        std::vector<unsigned long long> tt(4096 * 24), tn(4096 * 16);
        yh_x_trace_copy(tt.data(), 4096 * 24);
        yh_x_ntrace_copy(tn.data(), 4096 * 16);
        int G = 0; while (G < 4095 && tt[G * 24] != 0) ++G;
        unsigned long long t0 = ~0ull; for (int b = 0; b < G; ++b) t0 = std::min(t0, tt[b * 24]);
        auto show = [&](const char* name, std::vector<double> v) {
            if (v.empty()) return; std::sort(v.begin(), v.end());
            printf("  chain %-24s n=%3zu min %6.0f  p10 %6.0f  med %6.0f  p90 %6.0f  max %6.0f ns\n", name, v.size(), v[0], v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back());
        };
        std::vector<double> a, b2;
        for (int b = 0; b < G; ++b) { a.push_back((double)(tt[b * 24] - t0)); b2.push_back((double)(tt[b * 24 + 3] - t0)); }
        show("train CTA start", a); show("train tile loop done", b2);
        const char* nm[16] = {"nms start", "init done", "A issued+arrived", "-", "B decode done", "-", "-", "D pairs done", "D fixed point done", "-", "-", "-", "D done", "E emit done", "-", "-"};
        for (int sl = 0; sl < 14; ++sl) {
            if (nm[sl][0] == '-') continue;
            std::vector<double> v; for (int b = 0; b < N; ++b) v.push_back((double)((long long)tn[b * 16 + sl] - (long long)t0));
            show(nm[sl], v);
        }
    }
#endif
    float loss; CK(cudaMemcpy(&loss, sets[0].loss, 4, cudaMemcpyDeviceToHost));
    std::vector<int> kc(N); CK(cudaMemcpy(kc.data(), sets[0].kcnt, N * 4, cudaMemcpyDeviceToHost));
    long kept = 0; for (int v : kc) kept += v;
    printf("M=%d loss=%.6f kept=%ld (%.1f per image)\n", M, loss, kept, (double)kept / N);
    return 0;
}
