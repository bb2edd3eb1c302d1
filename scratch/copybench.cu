// Scratch micro-benchmark (not part of the product): what does a 21.6 MB -> 21.6 MB streaming
// pass cost on B200 under the bench's timing regime (CUDA graph of 8 launches over 8 rotating
// buffer pairs, replayed)?  Variants: LDG/STG grid-stride copy, per-warp TMA pipelines, read-only,
// write-only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <functional>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// A: grid-stride float4 copy
template <int UNROLL>
__global__ void copy_ldg(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(dst + i + u * stride, v[u]);
    }
    for (; i < n4; i += stride) dst[i] = src[i];
}

// contiguous-per-CTA float4 copy (each CTA owns a contiguous slab)
template <int UNROLL>
__global__ void copy_slab(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    const size_t b0 = n4 * blockIdx.x / gridDim.x, b1 = n4 * (blockIdx.x + 1) / gridDim.x;
    size_t i = b0 + threadIdx.x;
    for (; i + (UNROLL - 1) * blockDim.x < b1; i += UNROLL * blockDim.x) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * blockDim.x);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(dst + i + u * blockDim.x, v[u]);
    }
    for (; i < b1; i += blockDim.x) dst[i] = src[i];
}

__global__ void read_only(const float4* __restrict__ src, float* out, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) acc += src[i].x;
    if (acc == 1234.5678f) out[0] = acc;
}

__global__ void write_only(float4* __restrict__ dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(dst + i, make_float4(0.f, 1.f, 2.f, 3.f));
}

// B: per-warp TMA pipelines.  Each CTA owns a contiguous slab, cut into chunks of CH bytes dealt
// round-robin to its warps; each warp: IN stages of loads in flight, store straight from the
// landed stage (optionally through a smem->smem copy into an out stage, like the real kernel).
template <int IN, bool STAGE_OUT>
__global__ void copy_tma_warp(const char* __restrict__ src, char* __restrict__ dst, size_t bytes, int CH) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int OUT = STAGE_OUT ? 2 : 0;
    unsigned char* in_base = smem + (size_t)warp * (IN + OUT) * CH;
    unsigned char* out_base = in_base + (size_t)IN * CH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)W * (IN + OUT) * CH) + warp * IN;
    const size_t nchunks_total = (bytes + CH - 1) / CH;
    const size_t c0 = nchunks_total * blockIdx.x / gridDim.x, c1 = nchunks_total * (blockIdx.x + 1) / gridDim.x;
    const int nch = (int)(c1 - c0);
    const int my_n = nch > warp ? (nch - warp + W - 1) / W : 0;
    auto issue = [&](int k) {
        const size_t off = (c0 + warp + (size_t)k * W) * CH;
        const uint32_t b = (uint32_t)min((size_t)CH, bytes - off);
        mbar_expect_tx(&bars[k % IN], b);
        bulk_load(in_base + (size_t)(k % IN) * CH, src + off, b, &bars[k % IN]);
    };
    if (lane == 0) {
        for (int s = 0; s < IN; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int k = 0; k < min(IN, my_n); ++k) issue(k);
    }
    __syncwarp();
    for (int k = 0; k < my_n; ++k) {
        const size_t off = (c0 + warp + (size_t)k * W) * CH;
        const uint32_t b = (uint32_t)min((size_t)CH, bytes - off);
        mbar_wait(&bars[k % IN], (k / IN) & 1);
        unsigned char* in = in_base + (size_t)(k % IN) * CH;
        if (STAGE_OUT) {
            unsigned char* out = out_base + (size_t)(k % 2) * CH;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            for (uint32_t i = lane * 16; i < b; i += 512) *reinterpret_cast<float4*>(out + i) = *reinterpret_cast<const float4*>(in + i);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                bulk_store(dst + off, out, b);
                bulk_commit();
                if (k + IN < my_n) issue(k + IN);
            }
        } else {
            if (lane == 0) {
                bulk_store(dst + off, in, b);
                bulk_commit();
                if (k + IN < my_n) { bulk_wait_read<0>(); issue(k + IN); }
            }
        }
    }
    if (lane == 0) bulk_wait<0>();
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)256 * 84500;  // 21.632 MB
    const int R = 8;
    std::vector<char*> src(R), dst(R);
    for (int i = 0; i < R; ++i) { CK(cudaMalloc(&src[i], bytes + 256)); CK(cudaMalloc(&dst[i], bytes + 256)); CK(cudaMemset(src[i], 1, bytes)); }
    float* dummy; CK(cudaMalloc(&dummy, 256));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const size_t n4 = bytes / 16;

    auto run = [&](const char* name, double moved, std::function<void(int)> launch) {
        for (int i = 0; i < R; ++i) launch(i);
        CK(cudaStreamSynchronize(st));
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < R; ++i) launch(i);
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
        printf("%-44s %7.2f us  %7.0f GB/s\n", name, us, moved / us / 1e3);
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    };

    char name[128];
    for (int k : {1, 2, 4, 8, 16}) {
        snprintf(name, sizeof name, "copy_ldg<4> grid=148x%d x256", k);
        run(name, 2.0 * bytes, [&](int i) { copy_ldg<4><<<148 * k, 256, 0, st>>>((const float4*)src[i], (float4*)dst[i], n4); });
    }
    for (int k : {1, 2, 4}) {
        snprintf(name, sizeof name, "copy_ldg<8> grid=148x%d x512", k);
        run(name, 2.0 * bytes, [&](int i) { copy_ldg<8><<<148 * k, 512, 0, st>>>((const float4*)src[i], (float4*)dst[i], n4); });
    }
    for (int k : {1, 2, 4}) {
        snprintf(name, sizeof name, "copy_slab<8> grid=148x%d x512", k);
        run(name, 2.0 * bytes, [&](int i) { copy_slab<8><<<148 * k, 512, 0, st>>>((const float4*)src[i], (float4*)dst[i], n4); });
    }
    run("read_only grid=148x4 x512", 1.0 * bytes, [&](int i) { read_only<<<148 * 4, 512, 0, st>>>((const float4*)src[i], dummy, n4); });
    run("read_only grid=148x8 x256", 1.0 * bytes, [&](int i) { read_only<<<148 * 8, 256, 0, st>>>((const float4*)src[i], dummy, n4); });
    run("write_only grid=148x4 x512", 1.0 * bytes, [&](int i) { write_only<<<148 * 4, 512, 0, st>>>((float4*)dst[i], n4); });
    run("empty-ish kernel (write 1 CTA)", 1.0, [&](int i) { write_only<<<1, 32, 0, st>>>((float4*)dst[i], 32); });

    struct Cfg { int warps, ch, grid_mul; };
    for (Cfg c : {Cfg{8, 4096, 1}, Cfg{8, 8192, 1}, Cfg{16, 4096, 1}, Cfg{4, 8192, 2}, Cfg{8, 4096, 2}, Cfg{4, 16384, 1}, Cfg{8, 2048, 2}, Cfg{2, 16384, 2}}) {
        {
            const size_t smem = (size_t)c.warps * 3 * c.ch + c.warps * 3 * 8 + 16;
            CK(cudaFuncSetAttribute(copy_tma_warp<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            snprintf(name, sizeof name, "tma_warp<3,direct> w=%d ch=%d grid=148x%d", c.warps, c.ch, c.grid_mul);
            run(name, 2.0 * bytes, [&](int i) { copy_tma_warp<3, false><<<148 * c.grid_mul, c.warps * 32, smem, st>>>(src[i], dst[i], bytes, c.ch); });
        }
        {
            const size_t smem = (size_t)c.warps * 5 * c.ch + c.warps * 3 * 8 + 16;
            if (smem * c.grid_mul > 220 * 1024) continue;
            CK(cudaFuncSetAttribute(copy_tma_warp<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            snprintf(name, sizeof name, "tma_warp<3,staged>  w=%d ch=%d grid=148x%d", c.warps, c.ch, c.grid_mul);
            run(name, 2.0 * bytes, [&](int i) { copy_tma_warp<3, true><<<148 * c.grid_mul, c.warps * 32, smem, st>>>(src[i], dst[i], bytes, c.ch); });
        }
    }
    // sanity
    CK(cudaMemset(dst[0], 0, bytes));
    {
        const size_t smem = (size_t)8 * 5 * 4096 + 8 * 3 * 8 + 16;
        copy_tma_warp<3, true><<<148, 256, smem, st>>>(src[0], dst[0], bytes, 4096);
        CK(cudaStreamSynchronize(st));
        std::vector<char> h(bytes);
        CK(cudaMemcpy(h.data(), dst[0], bytes, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (size_t i = 0; i < bytes; ++i) bad += h[i] != 1;
        printf("sanity: %zu bad bytes\n", bad);
    }
    return 0;
}
