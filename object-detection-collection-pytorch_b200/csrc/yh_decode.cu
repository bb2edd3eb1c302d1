// predict(): the six decoded tensors from a head tensor.
// Replaces reference models/yolov2.py:469-649 and models/yolov1.py:250-437.
//
// One CTA per chunk of cells (a multiple of 4 cells, <= 16 KB, so that the chunk start is
// 16-byte aligned): the chunk is pulled into shared memory with one 1-D TMA bulk copy, box
// outputs are produced by one thread per predictor (float2/float4 stores, consecutive
// predictors -> coalesced), and the two [.., C] class tensors by a flat sweep over
// (predictor, class) so that their stores are coalesced too.  Several CTAs share an SM, which
// is what overlaps the loads of one chunk with the stores of another.
#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunkBytes = 16 * 1024;

struct DecodeParams {
    YhGeom g;
    const float* y;
    float2* sig_txty;
    float2* wh_act;
    float4* bbox;
    float* conf;
    float* cls_prob;
    float* cls_spec;
    long long total_cells;
    int cells_per_chunk;
    int tma_in;
};

// TV/TA/TC != 0 fix version / boxes per cell / classes at compile time (the divisions by C and A of the
// coalesced sweeps become multiplications); 0 keeps them as run-time values from the geometry.
template <int TV, int TA, int TC>
__global__ void __launch_bounds__(kThreads) yh_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const YhGeom& g = p.g;
    const int tid = threadIdx.x;
    const int version = TV ? TV : g.version;
    const int A = TA ? TA : g.a, C = TC ? TC : g.c;
    const int bs = version == 2 ? 5 + C : 5, cf = version == 2 ? A * (5 + C) : 5 * A + C;
    const int chunk_floats = p.cells_per_chunk * cf;
    float* in = reinterpret_cast<float*>(smem_raw);
    float* s_conf = in + chunk_floats;                       // [rows]
    float* s_mx = s_conf + p.cells_per_chunk * A;            // [groups] softmax max
    float* s_inv = s_mx + p.cells_per_chunk * A;             // [groups] 1/sum... kept as the sum
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_inv + p.cells_per_chunk * A);

    const long long cell0 = (long long)blockIdx.x * p.cells_per_chunk;
    const int ncell = (int)min((long long)p.cells_per_chunk, p.total_cells - cell0);
    const int nfl = ncell * cf, nfl16 = nfl & ~3;
    const float* ysrc = p.y + cell0 * cf;

    if (p.tma_in) {
        if (tid == 0) {
            yh_mbar_init(bar, 1);
            yh_mbar_fence_init();
            if (nfl16) {
                yh_mbar_expect_tx(bar, (uint32_t)nfl16 * 4u);
                yh_bulk_load(in, ysrc, (uint32_t)nfl16 * 4u, bar);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(yh_smem_u32(bar)) : "memory");
            }
        }
        if (tid < nfl - nfl16) in[nfl16 + tid] = __ldg(ysrc + nfl16 + tid);
        __syncthreads();  // barrier init visible before anyone polls it
        yh_mbar_wait(bar, 0);
    } else {
        for (int i = tid; i < nfl; i += kThreads) in[i] = __ldg(ysrc + i);
    }
    __syncthreads();

    const int cells = g.cells;
    const int nrows = ncell * A;
    const long long row0 = cell0 * A;  // global predictor index of the chunk's first row
    // class groups: one softmax per predictor (v2) or per cell (v1)
    const int ngroups = version == 2 ? nrows : ncell;

    for (int u = tid; u < nrows; u += kThreads) {
        const int lcell = u / A, a = u - lcell * A;
        const long long gcell = cell0 + lcell;
        const int icell = (int)(gcell % cells);
        const int cy = icell / g.s_w, cx = icell - cy * g.s_w;
        const float* bp = in + lcell * cf + a * bs;
        const float sx = yh_sigmoid(bp[0]), sy = yh_sigmoid(bp[1]);
        float wa, ha;
        if (version == 2) {
            wa = expf(bp[2]);
            ha = expf(bp[3]);
        } else {
            wa = yh_sigmoid(bp[2]);
            ha = yh_sigmoid(bp[3]);
        }
        const float conf = yh_sigmoid(bp[4]);
        s_conf[u] = conf;
        const YhBox b = yh_decode_box(sx, sy, wa, ha, g.pw[a], g.ph[a], cx, cy, g.gw, g.gh);
        const long long r = row0 + u;
        if (p.sig_txty) p.sig_txty[r] = make_float2(sx, sy);
        if (p.wh_act) p.wh_act[r] = make_float2(wa, ha);
        if (p.bbox) p.bbox[r] = make_float4(b.x1, b.y1, b.x2, b.y2);
        if (p.conf) p.conf[r] = conf;
    }
    if (!p.cls_prob && !p.cls_spec) return;

    // softmax statistics, one thread per class group; the exponentials replace the logits in the staged chunk
    // (nothing reads the class logits again), so the coalesced sweep below only divides
    for (int q = tid; q < ngroups; q += kThreads) {
        float* cl = version == 2 ? in + q * bs + 5 : in + q * cf + 5 * A;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, cl[c]);
        float se = 0.f;
        for (int c = 0; c < C; ++c) {
            const float e = expf(cl[c] - mx);
            cl[c] = e;
            se += e;
        }
        s_inv[q] = se;
    }
    __syncthreads();

    if (version == 2) {
        const int total = nrows * C;
        const long long o0 = row0 * C;
        for (int i = tid; i < total; i += kThreads) {
            const int u = i / C, c = i - u * C;
            const float pc = __fdiv_rn(in[u * bs + 5 + c], s_inv[u]);
            if (p.cls_prob) p.cls_prob[o0 + i] = pc;
            if (p.cls_spec) p.cls_spec[o0 + i] = __fmul_rn(pc, s_conf[u]);
        }
    } else {
        if (p.cls_prob) {
            const int total = ncell * C;
            const long long o0 = cell0 * C;
            for (int i = tid; i < total; i += kThreads) {
                const int q = i / C, c = i - q * C;
                p.cls_prob[o0 + i] = __fdiv_rn(in[q * cf + 5 * A + c], s_inv[q]);
            }
        }
        if (p.cls_spec) {
            const int total = nrows * C;
            const long long o0 = row0 * C;
            for (int i = tid; i < total; i += kThreads) {
                const int u = i / C, c = i - u * C;
                const int q = u / A;
                const float pc = __fdiv_rn(in[q * cf + 5 * A + c], s_inv[q]);
                p.cls_spec[o0 + i] = __fmul_rn(pc, s_conf[u]);
            }
        }
    }
}

template <int TV, int TA, int TC>
int launch_decode(const DecodeParams& p, unsigned grid, size_t smem, void* stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        int rc = yh_check_cuda(cudaFuncSetAttribute(yh_decode_kernel<TV, TA, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                               "cudaFuncSetAttribute(decode)");
        if (rc) return rc;
        configured[dev] = smem;
    }
    yh_decode_kernel<TV, TA, TC><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_decode launch");
}

int decode_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
                const float* anchors_wh_host, float img_h, float img_w, float* sig_txty,
                float* wh_act, float* bbox, float* conf, float* cls_prob, float* cls_spec,
                void* stream) {
    DecodeParams p;
    int rc = yh_make_geom(&p.g, version, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w);
    if (rc) return rc;
    YH_REQUIRE(y, YH_ERR_INVALID, "y is NULL");
    YH_REQUIRE(((uintptr_t)y & 3) == 0, YH_ERR_INVALID, "y must be 4-byte aligned");
    YH_REQUIRE(((uintptr_t)sig_txty & 7) == 0 && ((uintptr_t)wh_act & 7) == 0 && ((uintptr_t)bbox & 15) == 0,
               YH_ERR_INVALID, "sig_txty/wh_act need 8-byte and bbox 16-byte alignment");
    p.y = y;
    p.sig_txty = reinterpret_cast<float2*>(sig_txty);
    p.wh_act = reinterpret_cast<float2*>(wh_act);
    p.bbox = reinterpret_cast<float4*>(bbox);
    p.conf = conf; p.cls_prob = cls_prob; p.cls_spec = cls_spec;
    p.total_cells = (long long)n * p.g.cells;
    const int cf = p.g.cell_floats;
    int qpc = kChunkBytes / (16 * cf);
    if (qpc < 1) qpc = 1;
    p.cells_per_chunk = qpc * 4;
    p.tma_in = ((uintptr_t)y & 15) == 0;
    const size_t smem = (size_t)p.cells_per_chunk * cf * 4 + (size_t)3 * p.cells_per_chunk * a * 4 + 16;
    YH_REQUIRE(smem <= 227 * 1024, YH_ERR_UNSUPPORTED, "cell too wide for shared memory (%d floats)", cf);
    const long long grid = (p.total_cells + p.cells_per_chunk - 1) / p.cells_per_chunk;
    YH_REQUIRE(grid < (1ll << 31), YH_ERR_UNSUPPORTED, "too many chunks");
    // compile-time geometries for the shapes the reference uses (VOC: YOLOv2 5 anchors x 20 classes, YOLOv1 B=2, C=20)
    if (version == 2 && a == 5 && c == 20) return launch_decode<2, 5, 20>(p, (unsigned)grid, smem, stream);
    if (version == 1 && a == 2 && c == 20) return launch_decode<1, 2, 20>(p, (unsigned)grid, smem, stream);
    return launch_decode<0, 0, 0>(p, (unsigned)grid, smem, stream);
}

}  // namespace

extern "C" {

int yh_v2_decode(const float* y, int n, int s_h, int s_w, int a, int c, const float* anchors_wh_host,
                 float img_h, float img_w, float* sig_txty, float* wh_act, float* bbox, float* conf,
                 float* cls_prob, float* cls_spec, void* stream) {
    return decode_impl(2, y, n, s_h, s_w, a, c, anchors_wh_host, img_h, img_w, sig_txty, wh_act, bbox,
                       conf, cls_prob, cls_spec, stream);
}

int yh_v1_decode(const float* y, int n, int s_h, int s_w, int b, int c, float img_h, float img_w,
                 float* sig_txty, float* wh_act, float* bbox, float* conf, float* cls_prob,
                 float* cls_spec, void* stream) {
    return decode_impl(1, y, n, s_h, s_w, b, c, nullptr, img_h, img_w, sig_txty, wh_act, bbox, conf,
                       cls_prob, cls_spec, stream);
}

}  // extern "C"
