// Fused multi-tensor SGD step: every parameter tensor of the model in ONE launch.
//
// Replaces the optimizer step of run_one_epoch (reference models/yolov2.py:1253-1272,
// models/yolov1.py run_one_epoch likewise):
//     opt = SGD(self.parameters(), lr=..., momentum=0.9, weight_decay=5e-4)   # a NEW optimizer every iteration
//     opt.zero_grad(); loss.backward(); opt.step()
// torch.optim.SGD.step per tensor:  d = g + weight_decay * p;  buf = d on an optimizer's first step, else
// buf = momentum * buf + d;  p = p - lr * buf.  Because the reference builds a new optimizer in every
// iteration, every step is a first step: p -= lr * (g + wd * p), the momentum never accumulates (SURVEY
// B-11).  YH_SGD_FRESH_MOMENTUM reproduces exactly that; without the flag the buffers persist (the caller
// keeps them), which is what the reference's hyper-parameters suggest was intended -- a change of training
// semantics, hence an explicit flag.
//
// HBM-bound elementwise work (12 bytes per parameter in the reference's mode: read p and g, write p;
// 20 with persistent buffers).  The tensors are cut into chunks of YH_SGD_CHUNK elements, listed in a
// table the caller builds once (yh_sgd_plan) and keeps on the device; a persistent grid walks the
// table, 16-byte accesses where the three pointers allow it.
#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 8;

struct SgdParams {
    const YhSgdChunk* chunks;
    long long n_chunks;
    float lr, momentum, wd;
    int fresh;
};

__device__ __forceinline__ float sgd_one(float p, float g, float* b, const SgdParams& s, bool has_buf) {
    float d = s.wd != 0.f ? fmaf(s.wd, p, g) : g;          // grad.add(param, alpha=weight_decay)
    if (s.momentum != 0.f) {
        if (!s.fresh && has_buf) d = __fadd_rn(__fmul_rn(s.momentum, *b), d);  // buf.mul_(momentum).add_(grad)
        *b = d;
    }
    return fmaf(-s.lr, d, p);                               // param.add_(buf, alpha=-lr)
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm) yh_sgd_kernel(const SgdParams s) {
    yh_grid_dependency_wait();
    yh_grid_launch_dependents();
    for (long long c = blockIdx.x; c < s.n_chunks; c += gridDim.x) {
        const YhSgdChunk ch = s.chunks[c];
        const int n = ch.n;
        const bool has_buf = ch.buf != nullptr;
        const bool vec = (((uintptr_t)ch.p | (uintptr_t)ch.g | (uintptr_t)ch.buf) & 15) == 0;
        const int n4 = vec ? n >> 2 : 0;
        float4* p4 = reinterpret_cast<float4*>(ch.p);
        const float4* g4 = reinterpret_cast<const float4*>(ch.g);
        float4* b4 = reinterpret_cast<float4*>(ch.buf);
#pragma unroll 4
        for (int i = threadIdx.x; i < n4; i += kThreads) {
            float4 p = p4[i];
            const float4 g = __ldcs(g4 + i);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_buf && !s.fresh && s.momentum != 0.f) b = b4[i];
            p.x = sgd_one(p.x, g.x, &b.x, s, has_buf);
            p.y = sgd_one(p.y, g.y, &b.y, s, has_buf);
            p.z = sgd_one(p.z, g.z, &b.z, s, has_buf);
            p.w = sgd_one(p.w, g.w, &b.w, s, has_buf);
            p4[i] = p;
            if (has_buf && s.momentum != 0.f) b4[i] = b;
        }
        for (int i = 4 * n4 + threadIdx.x; i < n; i += kThreads) {
            float b = (has_buf && !s.fresh && s.momentum != 0.f) ? ch.buf[i] : 0.f;
            ch.p[i] = sgd_one(ch.p[i], ch.g[i], &b, s, has_buf);
            if (has_buf && s.momentum != 0.f) ch.buf[i] = b;
        }
    }
}

}  // namespace

extern "C" {

int64_t yh_sgd_chunk_count(const int64_t* sizes_host, int n_tensors) {
    if (!sizes_host || n_tensors < 0) return -1;
    int64_t c = 0;
    for (int t = 0; t < n_tensors; ++t) {
        if (sizes_host[t] < 0) return -1;
        c += (sizes_host[t] + YH_SGD_CHUNK - 1) / YH_SGD_CHUNK;
    }
    return c;
}

int yh_sgd_plan(const uint64_t* p_ptrs_host, const uint64_t* g_ptrs_host, const uint64_t* buf_ptrs_host,
                const int64_t* sizes_host, int n_tensors, YhSgdChunk* chunks_host, int64_t n_chunks) {
    YH_REQUIRE(p_ptrs_host && g_ptrs_host && sizes_host && (chunks_host || n_chunks == 0), YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(yh_sgd_chunk_count(sizes_host, n_tensors) == n_chunks, YH_ERR_INVALID, "n_chunks does not match the sizes");
    int64_t c = 0;
    for (int t = 0; t < n_tensors; ++t) {
        YH_REQUIRE(sizes_host[t] == 0 || (p_ptrs_host[t] && g_ptrs_host[t]), YH_ERR_INVALID, "tensor %d: null parameter or gradient", t);
        YH_REQUIRE(((p_ptrs_host[t] | g_ptrs_host[t] | (buf_ptrs_host ? buf_ptrs_host[t] : 0)) & 3) == 0, YH_ERR_INVALID,
                   "tensor %d: pointers must be 4-byte aligned", t);
        for (int64_t o = 0; o < sizes_host[t]; o += YH_SGD_CHUNK, ++c) {
            YhSgdChunk& ch = chunks_host[c];
            ch.p = reinterpret_cast<float*>(p_ptrs_host[t]) + o;
            ch.g = reinterpret_cast<const float*>(g_ptrs_host[t]) + o;
            ch.buf = (buf_ptrs_host && buf_ptrs_host[t]) ? reinterpret_cast<float*>(buf_ptrs_host[t]) + o : nullptr;
            const int64_t left = sizes_host[t] - o;
            ch.n = (int32_t)(left < YH_SGD_CHUNK ? left : YH_SGD_CHUNK);
            ch.reserved = 0;
        }
    }
    return 0;
}

int yh_sgd_step(const YhSgdChunk* chunks_dev, int64_t n_chunks, float lr, float momentum, float weight_decay,
                int flags, void* stream) {
    YH_REQUIRE(n_chunks >= 0 && (chunks_dev || n_chunks == 0), YH_ERR_INVALID, "bad chunk table");
    YH_REQUIRE(((uintptr_t)chunks_dev & 7) == 0, YH_ERR_INVALID, "chunk table must be 8-byte aligned");
    if (n_chunks == 0) return 0;
    SgdParams s;
    s.chunks = chunks_dev; s.n_chunks = n_chunks;
    s.lr = lr; s.momentum = momentum; s.wd = weight_decay;
    s.fresh = (flags & YH_SGD_FRESH_MOMENTUM) != 0;
    long long grid = (long long)yh_sm_count() * kCtasPerSm;
    if (grid > n_chunks) grid = n_chunks;
    return yh_check_cuda(yh_launch_pdl(yh_sgd_kernel, dim3((unsigned)grid), dim3(kThreads), 0, (cudaStream_t)stream, s),
                         "yh_sgd launch");
}

}  // extern "C"
