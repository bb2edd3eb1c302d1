// Shared device/host helpers for libyolohead (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "yolohead.h"

#ifndef __CUDA_ARCH__
#define YH_HOST_ONLY 1
#endif

// ------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, never throws)
// ------------------------------------------------------------------------------------------
void yh_set_error(const char* fmt, ...);
int yh_check_cuda(cudaError_t e, const char* what);
int yh_sm_count();  // SM count of the current device (cached per device)

#define YH_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            yh_set_error(__VA_ARGS__);         \
            return (code);                     \
        }                                      \
    } while (0)

// Geometry of a head tensor, passed by value to every kernel.
struct YhGeom {
    int version;      // 1 or 2
    int n, s_h, s_w;  // images, grid
    int a, c;         // anchors (v2) / boxes per cell (v1), classes
    int cells;        // s_h * s_w
    int cell_floats;  // v2: a*(5+c); v1: 5*a+c
    int box_stride;   // floats between consecutive boxes of a cell: v2 5+c, v1 5
    int preds;        // cells * a   (predictors per image)
    float gw, gh;     // pixels per grid cell: float(double(W)/S_w), float(double(H)/S_h)
    float pw[YH_MAX_ANCHORS];  // v2 anchor priors; v1: float(S_w) for every box
    float ph[YH_MAX_ANCHORS];  //                   v1: float(S_h)
};

int yh_make_geom(YhGeom* g, int version, int n, int s_h, int s_w, int a, int c,
                 const float* anchors_wh_host, float img_h, float img_w);

// How the train head cuts the flattened batch of cells into tiles (yh_train.cu).
int yh_train_tiling(long long total_cells, int cell_floats, int* tile_cells, int* num_tiles, int* grid);

// sigmoid(t) >= conf_thre decided on the logit: below to_reject never, from to_accept on always, in between
// the sigmoid is evaluated (yh_nms.cu; the fused train head applies the same rule, so both list the same set).
void yh_conf_band(float conf_thre, float* to_reject, float* to_accept);

// Totals of the six fixed-point loss sums -> terms and loss (yh_finalize.cuh); with world > 1 the sums of
// all ranks are exchanged through peer memory first.
struct YhFinalParams {
    unsigned long long* acc;   // [8] the train workspace: six sums, non-finite flags, (unused)
    float* terms;
    float* loss;
    float lam[5];
    double inv_den[5];
    int rank, world;
    unsigned long long* peer[YH_MAX_RANKS];  // exchange buffers of all ranks mapped into this process
};
int yh_fill_exchange(YhFinalParams* f, const YhExchange* xch_host);
int yh_launch_finalize(const YhFinalParams& f, cudaStream_t stream);  // the one-warp finalize kernel (yh_train.cu)

// gradient coefficients of the five loss terms (yh_loss_coefs fills them and the finalize parameters' lam / inv_den)
struct YhLossCoef {
    float cxy, cwh, cconf, cno, ccls;
};
void yh_loss_coefs(const float* lambdas_host, int m_global, int preds, YhLossCoef* k, YhFinalParams* f);

// The train head behind every train entry point (yh_train.cu), and the post-process behind every post-process
// entry point (yh_nms.cu); the fused step (yh_v2_train_post) sequences the two.
int yh_train_impl(int version, const float* y, int n, int s_h, int s_w, int a, int c,
                  const float* anchors_wh_host, float img_h, float img_w, const YhGt* gt,
                  const int32_t* gt_off, int m_local, int m_global, const float* lambdas_host,
                  float* dy, float* terms, float* loss, int32_t* resp, float* iou_resp, void* ws,
                  size_t ws_bytes, void* stream, int late_wait, const YhExchange* xch_host);

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// --- arithmetic with the reference's rounding: every torch elementwise op rounds once, so the
// decision path (decode -> corners -> IoU) uses explicit round-to-nearest intrinsics, which
// nvcc never contracts into FMAs.
__device__ __forceinline__ float yh_sigmoid(float x) {
    // torch.sigmoid: 1 / (1 + exp(-x))
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

struct YhBox {
    float x1, y1, x2, y2;
};

// Box decode, reference models/yolov2.py:488-595 / models/yolov1.py:275-382.
// wa/ha are the activated sizes (v2: exp(t), v1: sigmoid(t)); pw/ph the multipliers.
__device__ __forceinline__ YhBox yh_decode_box(float sx, float sy, float wa, float ha, float pw,
                                               float ph, int cx, int cy, float gw, float gh) {
    const float bw = __fmul_rn(pw, wa);
    const float bh = __fmul_rn(ph, ha);
    const float bx = __fadd_rn(sx, (float)cx);
    const float by = __fadd_rn(sy, (float)cy);
    const float hw = __fmul_rn(bw, 0.5f);  // bw / 2 of the reference, bit for bit (power of two)
    const float hh = __fmul_rn(bh, 0.5f);
    YhBox b;
    b.x1 = __fmul_rn(__fsub_rn(bx, hw), gw);
    b.y1 = __fmul_rn(__fsub_rn(by, hh), gh);
    b.x2 = __fmul_rn(__fadd_rn(bx, hw), gw);
    b.y2 = __fmul_rn(__fadd_rn(by, hh), gh);
    return b;
}

// IoU, reference models/utils.py:47-63: inter / (union + 1e-6), first box = coord1.
__device__ __forceinline__ float yh_iou_xyxy(const YhBox& p, const YhBox& q) {
    const float iw = fmaxf(__fsub_rn(fminf(p.x2, q.x2), fmaxf(p.x1, q.x1)), 0.0f);
    const float ih = fmaxf(__fsub_rn(fminf(p.y2, q.y2), fmaxf(p.y1, q.y1)), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float a1 = __fmul_rn(__fsub_rn(p.x2, p.x1), __fsub_rn(p.y2, p.y1));
    const float a2 = __fmul_rn(__fsub_rn(q.x2, q.x1), __fsub_rn(q.y2, q.y1));
    const float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
    return __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));
}

__device__ __forceinline__ float yh_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float yh_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// --- mbarrier + 1-D TMA bulk copies (cp.async.bulk; SASS: UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t yh_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void yh_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(yh_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void yh_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void yh_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(yh_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void yh_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "YH_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra YH_DONE_%=;\n\t"
        "bra YH_WAIT_%=;\n\t"
        "YH_DONE_%=:\n\t"
        "}" ::"r"(yh_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool yh_mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(yh_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared, completion signalled on `bar` (bytes: multiple of 16, both 16-B aligned)
__device__ __forceinline__ void yh_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(yh_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(yh_smem_u32(bar))
        : "memory");
}
// global -> L2 only (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void yh_bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void yh_bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(yh_smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void yh_bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void yh_bulk_wait_read() {  // <= N groups still reading smem
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void yh_bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (before a bulk store)
__device__ __forceinline__ void yh_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// --- programmatic dependent launch (PDL): a kernel launched with yh_launch_pdl may have its CTAs
// scheduled while the previous kernel of the stream is still draining; it must not touch global
// memory before yh_grid_dependency_wait().  Both are no-ops for an ordinary launch.
__device__ __forceinline__ void yh_grid_dependency_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void yh_grid_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

#endif  // __CUDACC__

#ifdef __CUDACC__
// Launch with the programmatic-stream-serialization attribute (CUDA-graph capturable): the launch
// latency and the prologue of this kernel overlap the tail of the previous one.
template <typename... KArgs, typename... Args>
inline cudaError_t yh_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif
