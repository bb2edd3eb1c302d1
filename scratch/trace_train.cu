// Scratch: per-warp timeline of yh_train_kernel (clock64 at phase boundaries) on the headline shape.
#define YH_TRACE 1
#include "../object-detection-collection-pytorch_b200/csrc/yh_api.cu"
#include "../object-detection-collection-pytorch_b200/csrc/yh_train.cu"
#include <vector>
#include <algorithm>
#include <random>

int main() {
    const int N = 256, S = 13, A = 5, C = 20;
    const size_t nfl = (size_t)N * S * S * A * (5 + C);
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> hy(nfl);
    for (auto& v : hy) v = nd(rng);
    std::vector<YhGt> gt;
    std::vector<int> off(N + 1, 0);
    for (int n = 0; n < N; ++n) {
        const int k = 1 + rng() % 5;
        for (int j = 0; j < k; ++j) {
            YhGt r;
            r.img = n; r.cy = rng() % S; r.cx = rng() % S; r.cls = rng() % C;
            r.stx = 0.3f; r.sty = 0.6f; r.tw = 2.5f; r.th = 3.5f;
            r.x1 = r.cx * 32.f - 20; r.y1 = r.cy * 32.f - 30; r.x2 = r.cx * 32.f + 60; r.y2 = r.cy * 32.f + 70;
            gt.push_back(r);
        }
        off[n + 1] = (int)gt.size();
    }
    float *y, *dy, *terms, *loss; YhGt* dgt; int* doff; void* ws;
    cudaMalloc(&y, nfl * 4); cudaMalloc(&dy, nfl * 4); cudaMalloc(&terms, 64); cudaMalloc(&loss, 64);
    cudaMalloc(&dgt, gt.size() * sizeof(YhGt)); cudaMalloc(&doff, (N + 1) * 4);
    cudaMalloc(&ws, yh_train_workspace_bytes()); cudaMemset(ws, 0, yh_train_workspace_bytes());
    cudaMemcpy(y, hy.data(), nfl * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dgt, gt.data(), gt.size() * sizeof(YhGt), cudaMemcpyHostToDevice);
    cudaMemcpy(doff, off.data(), (N + 1) * 4, cudaMemcpyHostToDevice);
    const float anchors[10] = {1.3221f, 1.73145f, 3.19275f, 4.00944f, 5.05587f, 8.09892f, 9.47112f, 4.84053f, 11.2364f, 10.0071f};
    const float lam[5] = {5, 5, 1, 0.5f, 1};
    float* flush; cudaMalloc(&flush, 256 << 20);
    for (int it = 0; it < 4; ++it) {
        cudaMemset(flush, it, 256 << 20);  // push y / dy out of L2
        int rc = yh_v2_train(y, N, S, S, A, C, anchors, 416, 416, dgt, doff, (int)gt.size(), (int)gt.size(), lam, dy, terms,
                             loss, nullptr, nullptr, ws, yh_train_workspace_bytes(), nullptr);
        if (rc) { printf("error %d %s\n", rc, yh_last_error()); return 1; }
        cudaDeviceSynchronize();
    }
    std::vector<long long> tr(1024 * 8 * 64);
    cudaMemcpyFromSymbol(tr.data(), g_trace, tr.size() * 8);
    float hl; cudaMemcpy(&hl, loss, 4, cudaMemcpyDeviceToHost);
    printf("loss %f, %s\n", hl, cudaGetErrorString(cudaGetLastError()));
    // per-warp summary for a few CTAs and aggregate phase durations
    double sum[8] = {0}; int cnt[8] = {0};
    double tot_kernel = 0, tot_pro = 0, tot_loop = 0, tot_drain = 0, tot_tail = 0; int nw = 0;
    long long maxend = 0;
    for (int b = 0; b < 148; ++b)
        for (int w = 0; w < 8; ++w) {
            long long* t = &tr[((size_t)b * 8 + w) * 64];
            if (!t[0]) continue;
            ++nw;
            tot_pro += t[1] - t[0]; tot_loop += t[2] - t[1]; tot_drain += t[3] - t[2]; tot_tail += t[63] - t[3]; tot_kernel += t[63] - t[0];
            maxend = std::max(maxend, t[63] - t[0]);
            for (int k = 0; k < 11 && t[8 + 5 * k]; ++k) {
                long long prev = k == 0 ? t[1] : t[8 + 5 * (k - 1)];
                sum[0] += t[4 + 5 * k] - prev; sum[1] += t[5 + 5 * k] - t[4 + 5 * k]; sum[2] += t[6 + 5 * k] - t[5 + 5 * k];
                sum[3] += t[7 + 5 * k] - t[6 + 5 * k]; sum[4] += t[8 + 5 * k] - t[7 + 5 * k]; cnt[0]++;
            }
            if ((b == 77) && t[57]) printf("   record: act %lld  gather+iou %lld  argmax+channels %lld  max %lld  exp+sum %lld  scalars %lld  cls-writes %lld\n", t[51]-t[50], t[52]-t[51], t[53]-t[52], t[54]-t[53], t[55]-t[54], t[56]-t[55], t[57]-t[56]);
            if (b == 77) {
                printf("   prologue: issue %lld  tma %lld zero %lld  sync1 %lld  meta-arrive %lld  sync2 %lld  rest %lld\n", t[58] - t[0], t[49] - t[58], t[59] - t[49], t[60] - t[59], t[61] - t[60], t[62] - t[61], t[1] - t[62]);
                printf("cta %3d warp %d: pro %lld loop %lld drain %lld tail %lld | chunks:", b, w, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[63] - t[3]);
                for (int k = 0; k < 6 && t[8 + 5 * k]; ++k) printf(" [w%lld d%lld s%lld p%lld]", t[4 + 5 * k] - (k == 0 ? t[1] : t[8 + 5 * (k - 1)]), t[6 + 5 * k] - t[5 + 5 * k], t[7 + 5 * k] - t[6 + 5 * k], t[8 + 5 * k] - t[7 + 5 * k]);
                printf("\n");
            }
        }
    {
        std::vector<std::pair<long long, int>> ends;
        for (int b = 0; b < 148; ++b) {
            long long e = 0;
            for (int w = 0; w < 8; ++w) { long long* t = &tr[((size_t)b * 8 + w) * 64]; if (t[0]) e = std::max(e, t[3] - t[0]); }
            ends.push_back({e, b});
        }
        std::sort(ends.begin(), ends.end());
        printf("per-CTA time to last store drained (cycles): min %lld  p25 %lld  median %lld  p75 %lld  p90 %lld  max %lld (cta %d)\n",
               ends[0].first, ends[37].first, ends[74].first, ends[111].first, ends[133].first, ends[147].first, ends[147].second);
        for (int i = 140; i < 148; ++i) {
            int b = ends[i].second;
            printf("  slow cta %3d: %lld |", b, ends[i].first);
            for (int w = 0; w < 8; ++w) { long long* t = &tr[((size_t)b * 8 + w) * 64]; printf(" w%d pro %lld first-wait %lld loop %lld", w, t[1]-t[0], t[4]-t[1], t[2]-t[1]); }
            printf("\n");
            // the slowest warp of this CTA, chunk by chunk
            int sw = 0; long long best = 0;
            for (int w = 0; w < 8; ++w) { long long* t = &tr[((size_t)b * 8 + w) * 64]; if (t[2] - t[1] > best) { best = t[2] - t[1]; sw = w; } }
            long long* t = &tr[((size_t)b * 8 + sw) * 64];
            printf("     slowest warp %d:", sw);
            for (int k = 0; k < 9 && t[8 + 5 * k]; ++k) printf(" [w%lld z%lld d%lld s%lld p%lld]", t[4 + 5 * k] - (k == 0 ? t[1] : t[8 + 5 * (k - 1)]), t[5 + 5 * k] - t[4 + 5 * k], t[6 + 5 * k] - t[5 + 5 * k], t[7 + 5 * k] - t[6 + 5 * k], t[8 + 5 * k] - t[7 + 5 * k]);
            printf("\n");
        }
    }
    printf("warps %d  avg cycles: kernel %.0f (max %lld) prologue %.0f loop %.0f drain %.0f tail %.0f\n", nw, tot_kernel / nw, maxend, tot_pro / nw, tot_loop / nw, tot_drain / nw, tot_tail / nw);
    printf("per chunk avg cycles: wait-load %.0f  zero/waitread %.0f  dense %.0f  sparse %.0f  store+refill %.0f  (chunks %d)\n",
           sum[0] / cnt[0], sum[1] / cnt[0], sum[2] / cnt[0], sum[3] / cnt[0], sum[4] / cnt[0], cnt[0]);
    return 0;
}
