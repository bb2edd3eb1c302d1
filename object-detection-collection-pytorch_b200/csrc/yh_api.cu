// libyolohead: error plumbing, geometry, and the two small elementwise entry points.
#include <stdarg.h>
#include <string.h>

#include "yh_common.cuh"

static thread_local char g_err[512] = "";

void yh_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int yh_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return YH_OK;
    yh_set_error("%s: %s", what, cudaGetErrorString(e));
    return YH_ERR_CUDA;
}

int yh_sm_count() {
    // read-only after first use per device; a benign race writes the same value twice
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cache[dev] = v;
    }
    return cache[dev];
}

int yh_make_geom(YhGeom* g, int version, int n, int s_h, int s_w, int a, int c,
                 const float* anchors_wh_host, float img_h, float img_w) {
    YH_REQUIRE(version == 1 || version == 2, YH_ERR_INVALID, "version must be 1 or 2");
    YH_REQUIRE(n > 0 && s_h > 0 && s_w > 0, YH_ERR_INVALID, "n, s_h, s_w must be positive (n=%d s_h=%d s_w=%d)", n, s_h, s_w);
    YH_REQUIRE(a > 0 && a <= YH_MAX_ANCHORS, YH_ERR_INVALID, "anchors/boxes per cell must be in 1..%d (got %d)", YH_MAX_ANCHORS, a);
    YH_REQUIRE(c > 0 && c <= 4096, YH_ERR_INVALID, "classes must be in 1..4096 (got %d)", c);
    YH_REQUIRE(img_h > 0.f && img_w > 0.f, YH_ERR_INVALID, "image size must be positive");
    YH_REQUIRE(version == 1 || anchors_wh_host != nullptr, YH_ERR_INVALID, "anchors_wh_host is NULL");
    YH_REQUIRE((long long)n * s_h * s_w * a < (1ll << 31), YH_ERR_UNSUPPORTED, "more than 2^31 predictors");
    memset(g, 0, sizeof(*g));
    g->version = version;
    g->n = n; g->s_h = s_h; g->s_w = s_w; g->a = a; g->c = c;
    g->cells = s_h * s_w;
    g->cell_floats = version == 2 ? a * (5 + c) : 5 * a + c;
    g->box_stride = version == 2 ? 5 + c : 5;
    g->preds = g->cells * a;
    // `height / num_grid_cell_in_height` is a Python float (double) that torch then applies
    // as an fp32 scalar -- models/yolov2.py:589-595
    g->gw = (float)((double)img_w / (double)s_w);
    g->gh = (float)((double)img_h / (double)s_h);
    for (int i = 0; i < a; ++i) {
        if (version == 2) {
            g->pw[i] = anchors_wh_host[2 * i];
            g->ph[i] = anchors_wh_host[2 * i + 1];
        } else {
            g->pw[i] = (float)s_w;  // models/yolov1.py:298-299
            g->ph[i] = (float)s_h;
        }
    }
    return YH_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void yh_iou_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                              long long count, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float4 p = __ldg(b1 + i), q = __ldg(b2 + i);
        YhBox bp{p.x, p.y, p.z, p.w}, bq{q.x, q.y, q.z, q.w};
        out[i] = yh_iou_xyxy(bp, bq);
    }
}

// float64 variant: get_iou(..., numpy=True) on float64 arrays (models/utils.py:30-38, 52-63; evaluate_model feeds it
// float64 boxes, :250-252).  numpy rounds every elementwise operation once: explicit round-to-nearest intrinsics
// keep nvcc from contracting them into FMAs, so the bits are numpy's.
__global__ void yh_iou_f64_kernel(const double* __restrict__ b1, const double* __restrict__ b2,
                                  long long count, double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double x1 = b1[4 * i], y1 = b1[4 * i + 1], x2 = b1[4 * i + 2], y2 = b1[4 * i + 3];
        const double u1 = b2[4 * i], v1 = b2[4 * i + 1], u2 = b2[4 * i + 2], v2 = b2[4 * i + 3];
        // np.maximum / np.minimum propagate NaN, np.clip(a_min=0) as well
        const double ix1 = (x1 != x1 || u1 != u1) ? x1 + u1 : fmax(x1, u1);
        const double iy1 = (y1 != y1 || v1 != v1) ? y1 + v1 : fmax(y1, v1);
        const double ix2 = (x2 != x2 || u2 != u2) ? x2 + u2 : fmin(x2, u2);
        const double iy2 = (y2 != y2 || v2 != v2) ? y2 + v2 : fmin(y2, v2);
        const double dw = __dsub_rn(ix2, ix1), dh = __dsub_rn(iy2, iy1);
        const double iw = dw != dw ? dw : fmax(dw, 0.0), ih = dh != dh ? dh : fmax(dh, 0.0);
        const double inter = __dmul_rn(iw, ih);
        const double a1 = __dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1));
        const double a2 = __dmul_rn(__dsub_rn(u2, u1), __dsub_rn(v2, v1));
        const double uni = __dsub_rn(__dadd_rn(a1, a2), inter);
        out[i] = __ddiv_rn(inter, __dadd_rn(uni, 1e-6));
    }
}

__global__ void yh_scale_kernel(float* __restrict__ x, long long count, const float* __restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) x[i] *= s;
}

extern "C" {

int yh_abi_version(void) { return YH_ABI_VERSION; }
const char* yh_last_error(void) { return g_err; }

int yh_iou(const float* boxes1, const float* boxes2, int64_t count, float* iou, void* stream) {
    YH_REQUIRE(count >= 0, YH_ERR_INVALID, "count < 0");
    if (count == 0) return YH_OK;
    YH_REQUIRE(boxes1 && boxes2 && iou, YH_ERR_INVALID, "null pointer");
    YH_REQUIRE(((uintptr_t)boxes1 & 15) == 0 && ((uintptr_t)boxes2 & 15) == 0, YH_ERR_INVALID,
               "box arrays must be 16-byte aligned");
    const int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    const long long cap = (long long)yh_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    yh_iou_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const float4*)boxes1, (const float4*)boxes2, (long long)count, iou);
    return yh_check_cuda(cudaGetLastError(), "yh_iou launch");
}

int yh_iou_f64(const double* boxes1, const double* boxes2, int64_t count, double* iou, void* stream) {
    YH_REQUIRE(count >= 0, YH_ERR_INVALID, "count < 0");
    if (count == 0) return YH_OK;
    YH_REQUIRE(boxes1 && boxes2 && iou, YH_ERR_INVALID, "null pointer");
    YH_REQUIRE(((uintptr_t)boxes1 & 7) == 0 && ((uintptr_t)boxes2 & 7) == 0 && ((uintptr_t)iou & 7) == 0, YH_ERR_INVALID,
               "arrays must be 8-byte aligned");
    const int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    const long long cap = (long long)yh_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    yh_iou_f64_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(boxes1, boxes2, (long long)count, iou);
    return yh_check_cuda(cudaGetLastError(), "yh_iou_f64 launch");
}

int yh_scale_inplace(float* x, int64_t count, const float* scale_dev, void* stream) {
    YH_REQUIRE(count >= 0, YH_ERR_INVALID, "count < 0");
    if (count == 0) return YH_OK;
    YH_REQUIRE(x && scale_dev, YH_ERR_INVALID, "null pointer");
    const int threads = 256;
    long long blocks = (count + threads * 4 - 1) / (threads * 4);
    const long long cap = (long long)yh_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    yh_scale_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(x, (long long)count, scale_dev);
    return yh_check_cuda(cudaGetLastError(), "yh_scale_inplace launch");
}

}  // extern "C"
