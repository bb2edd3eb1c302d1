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
// 20 with persistent buffers).  The tensor list travels in the kernel's parameter space (pointers and
// sizes of up to kMaxTensors tensors per launch: no device-side table, nothing to upload or cache, CUDA-graph
// capturable); the tensors are cut into chunks of YH_SGD_CHUNK elements and a persistent grid walks the
// chunks, 16-byte accesses where the three pointers allow it.
#include "yh_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 8;
constexpr int kMaxTensors = 96;  // per launch: 96 * 28 bytes + chunk prefix = 3.1 KB of the 4 KB parameter space

struct SgdParams {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* buf[kMaxTensors];
    int chunk_start[kMaxTensors + 1];  // first chunk of tensor t; [n_tensors] = all chunks
    int n[kMaxTensors];
    int n_tensors;
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

__global__ void __launch_bounds__(kThreads, kCtasPerSm) yh_sgd_kernel(const __grid_constant__ SgdParams s) {
    yh_grid_dependency_wait();
    yh_grid_launch_dependents();
    const int n_chunks = s.chunk_start[s.n_tensors];
    int t = 0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        while (s.chunk_start[t + 1] <= c) ++t;  // (uniform over the CTA; c only grows)
        const int o = (c - s.chunk_start[t]) * YH_SGD_CHUNK;
        const int n = min(YH_SGD_CHUNK, s.n[t] - o);
        float* cp = s.p[t] + o;
        const float* cg = s.g[t] + o;
        float* cb = s.buf[t] ? s.buf[t] + o : nullptr;
        const bool has_buf = cb != nullptr;
        const bool vec = (((uintptr_t)cp | (uintptr_t)cg | (uintptr_t)cb) & 15) == 0;
        const int n4 = vec ? n >> 2 : 0;
        float4* p4 = reinterpret_cast<float4*>(cp);
        const float4* g4 = reinterpret_cast<const float4*>(cg);
        float4* b4 = reinterpret_cast<float4*>(cb);
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
            float b = (has_buf && !s.fresh && s.momentum != 0.f) ? cb[i] : 0.f;
            cp[i] = sgd_one(cp[i], cg[i], &b, s, has_buf);
            if (has_buf && s.momentum != 0.f) cb[i] = b;
        }
    }
}

}  // namespace

extern "C" {

int yh_sgd_step(const uint64_t* p_ptrs_host, const uint64_t* g_ptrs_host, const uint64_t* buf_ptrs_host,
                const int64_t* sizes_host, int n_tensors, float lr, float momentum, float weight_decay,
                int flags, void* stream) {
    YH_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (p_ptrs_host && g_ptrs_host && sizes_host)), YH_ERR_INVALID,
               "null pointer argument");
    for (int t = 0; t < n_tensors; ++t) {
        YH_REQUIRE(sizes_host[t] >= 0 && sizes_host[t] < (1ll << 31), YH_ERR_UNSUPPORTED, "tensor %d: size %lld out of range", t,
                   (long long)sizes_host[t]);
        YH_REQUIRE(sizes_host[t] == 0 || (p_ptrs_host[t] && g_ptrs_host[t]), YH_ERR_INVALID, "tensor %d: null parameter or gradient", t);
        YH_REQUIRE(((p_ptrs_host[t] | g_ptrs_host[t] | (buf_ptrs_host ? buf_ptrs_host[t] : 0)) & 3) == 0, YH_ERR_INVALID,
                   "tensor %d: pointers must be 4-byte aligned", t);
    }
    const long long slots = (long long)yh_sm_count() * kCtasPerSm;
    for (int t0 = 0; t0 < n_tensors;) {  // one launch per kMaxTensors tensors (a model's whole list, normally)
        SgdParams s;
        int k = 0;
        long long chunks = 0;
        s.chunk_start[0] = 0;
        for (; t0 < n_tensors && k < kMaxTensors; ++t0) {
            if (sizes_host[t0] == 0) continue;
            const long long c = (sizes_host[t0] + YH_SGD_CHUNK - 1) / YH_SGD_CHUNK;
            if (chunks + c >= (1ll << 31)) break;
            s.p[k] = reinterpret_cast<float*>(p_ptrs_host[t0]);
            s.g[k] = reinterpret_cast<const float*>(g_ptrs_host[t0]);
            s.buf[k] = buf_ptrs_host ? reinterpret_cast<float*>(buf_ptrs_host[t0]) : nullptr;
            s.n[k] = (int)sizes_host[t0];
            chunks += c;
            s.chunk_start[++k] = (int)chunks;
        }
        if (k == 0) continue;
        s.n_tensors = k;
        s.lr = lr; s.momentum = momentum; s.wd = weight_decay;
        s.fresh = (flags & YH_SGD_FRESH_MOMENTUM) != 0;
        const long long grid = chunks < slots ? chunks : slots;
        int rc = yh_check_cuda(yh_launch_pdl(yh_sgd_kernel, dim3((unsigned)grid), dim3(kThreads), 0, (cudaStream_t)stream, s),
                               "yh_sgd launch");
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
