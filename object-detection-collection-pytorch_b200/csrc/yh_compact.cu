// Dense per-box target grids -> compact ground-truth records + per-image CSR offsets.
//
// The reference hands get_loss one [S_h,S_w,.] grid set per box with a single cell filled in
// (collate_fn, reference models/yolov2.py:1440-1555, models/yolov1.py:1238-1355) plus the image
// id of every box; get_loss then maps boxes to batch positions with an argmax over id equality
// (models/yolov2.py:828-834).  This adaptor recovers the 12 scalars per box so that the fused
// head kernel can run behind the unchanged get_loss signature.  It is the only part of the
// path whose traffic scales with M*S*S (it has to look at obj_mask), and it is reported
// separately from the head kernel's roofline.
//
// Three small launches on the caller's stream:
//   rows   : one warp per box -- image lookup, find the set cell, gather, class index
//   offsets: one warp        -- histogram -> exclusive scan -> gt_off
//   scatter: one warp per image -- stable scatter into image order (boxes keep their order)
#include "yh_common.cuh"

namespace {

struct CompactParams {
    const float* sig_txty;
    const float* twth;
    const float* coord;
    const float* cls_tgt;
    const void* obj_mask;
    int obj_is_f64;
    const int64_t* x_img_id;
    const int64_t* bbox_img_id;
    int m, n, s_h, s_w, c;
    YhGt* tmp;         // [m] unsorted records
    int32_t* key;      // [m] image of every row
    int32_t* count;    // [n]
    YhGt* gt_out;
    int32_t* gt_off;   // [n+1]
    int32_t* status;   // [2]
};

__global__ void yh_compact_init(CompactParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n) p.count[i] = 0;
    if (i < 2) p.status[i] = 0;
}

__global__ void __launch_bounds__(256) yh_compact_rows(CompactParams p) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p.m) return;
    const int cells = p.s_h * p.s_w;

    // image of this box: first n with x_img_id[n] == bbox_img_id[j], else 0
    const int64_t id = __ldg(p.bbox_img_id + j);
    int img = 0x7fffffff;
    for (int base = 0; base < p.n; base += 32) {
        const int n = base + lane;
        const bool hit = n < p.n && __ldg(p.x_img_id + n) == id;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) { img = base + __ffs(bal) - 1; break; }
    }
    if (img == 0x7fffffff) img = 0;

    // the set cell of obj_mask[j]
    int first = 0x7fffffff, nset = 0;
    for (int base = 0; base < cells; base += 32) {
        const int q = base + lane;
        bool set = false;
        if (q < cells) {
            const size_t o = (size_t)j * cells + q;
            set = p.obj_is_f64 ? (__ldg(reinterpret_cast<const double*>(p.obj_mask) + o) != 0.0)
                               : (__ldg(reinterpret_cast<const float*>(p.obj_mask) + o) != 0.f);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, set);
        if (bal) {
            if (first == 0x7fffffff) first = base + __ffs(bal) - 1;
            nset += __popc(bal);
        }
    }
    const int cell = first == 0x7fffffff ? 0 : first;
    const size_t row = (size_t)j * cells + cell;

    // class index: the one-hot position of cls_tgt[j, cell, :]
    int cls = 0x7fffffff, ones = 0, other = 0;
    for (int base = 0; base < p.c; base += 32) {
        const int c = base + lane;
        const float v = c < p.c ? __ldg(p.cls_tgt + row * p.c + c) : 0.f;
        const unsigned b1 = __ballot_sync(0xffffffffu, v == 1.f);
        const unsigned b0 = __ballot_sync(0xffffffffu, v != 1.f && v != 0.f);
        if (b1 && cls == 0x7fffffff) cls = base + __ffs(b1) - 1;
        ones += __popc(b1);
        other += __popc(b0);
    }
    if (lane == 0) {
        YhGt g;
        g.img = img;
        g.cy = cell / p.s_w;
        g.cx = cell - g.cy * p.s_w;
        g.cls = cls == 0x7fffffff ? -1 : cls;
        const float2 t = __ldg(reinterpret_cast<const float2*>(p.sig_txty) + row);
        const float2 w = __ldg(reinterpret_cast<const float2*>(p.twth) + row);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.coord) + row);
        g.stx = t.x; g.sty = t.y; g.tw = w.x; g.th = w.y;
        g.x1 = b.x; g.y1 = b.y; g.x2 = b.z; g.y2 = b.w;
        p.tmp[j] = g;
        p.key[j] = img;
        atomicAdd(p.count + img, 1);
        if (nset != 1) atomicAdd(p.status + 0, 1);
        if (ones != 1 || other != 0) atomicAdd(p.status + 1, 1);
    }
}

__global__ void yh_compact_offsets(CompactParams p) {  // one warp
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < p.n; base += 32) {
        const int n = base + lane;
        const int v = n < p.n ? p.count[n] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (n < p.n) p.gt_off[n] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) p.gt_off[p.n] = carry;
}

__global__ void __launch_bounds__(256) yh_compact_scatter(CompactParams p) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= p.n) return;
    const int lo = p.gt_off[n], hi = p.gt_off[n + 1];
    if (hi == lo) return;
    int pos = lo;
    for (int base = 0; base < p.m && pos < hi; base += 32) {
        const int j = base + lane;
        const bool hit = j < p.m && __ldg(p.key + j) == n;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int dst = pos + __popc(bal & ((1u << lane) - 1u));
            const int4* s = reinterpret_cast<const int4*>(p.tmp + j);
            int4* d = reinterpret_cast<int4*>(p.gt_out + dst);
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        }
        pos += __popc(bal);
    }
}

}  // namespace

extern "C" {

size_t yh_compact_workspace_bytes(int m, int n) {
    if (m < 0) m = 0;
    if (n < 0) n = 0;
    return (size_t)m * sizeof(YhGt) + ((size_t)m + (size_t)n) * 4 + 64;
}

int yh_compact_targets(const float* sig_txty, const float* twth, const float* coord,
                       const float* cls_tgt, const void* obj_mask, int obj_is_f64,
                       const int64_t* x_img_id, const int64_t* bbox_img_id, int m, int n, int s_h,
                       int s_w, int c, YhGt* gt_out, int32_t* gt_off_out, int32_t* status, void* ws,
                       size_t ws_bytes, void* stream) {
    YH_REQUIRE(m > 0, YH_ERR_EMPTY, "no ground-truth boxes (m=%d)", m);
    YH_REQUIRE(n > 0 && s_h > 0 && s_w > 0 && c > 0, YH_ERR_INVALID, "bad sizes");
    YH_REQUIRE(sig_txty && twth && coord && cls_tgt && obj_mask && x_img_id && bbox_img_id && gt_out &&
                   gt_off_out && status && ws,
               YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(ws_bytes >= yh_compact_workspace_bytes(m, n), YH_ERR_WORKSPACE, "workspace too small");
    YH_REQUIRE(((uintptr_t)ws & 15) == 0 && ((uintptr_t)gt_out & 15) == 0 && ((uintptr_t)coord & 15) == 0 &&
                   ((uintptr_t)sig_txty & 7) == 0 && ((uintptr_t)twth & 7) == 0,
               YH_ERR_INVALID, "misaligned pointer");
    CompactParams p;
    p.sig_txty = sig_txty; p.twth = twth; p.coord = coord; p.cls_tgt = cls_tgt;
    p.obj_mask = obj_mask; p.obj_is_f64 = obj_is_f64;
    p.x_img_id = x_img_id; p.bbox_img_id = bbox_img_id;
    p.m = m; p.n = n; p.s_h = s_h; p.s_w = s_w; p.c = c;
    p.tmp = reinterpret_cast<YhGt*>(ws);
    p.key = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + (size_t)m * sizeof(YhGt));
    p.count = p.key + m;
    p.gt_out = gt_out; p.gt_off = gt_off_out; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    const int wpb = 8;
    yh_compact_init<<<(max(n, 2) + 255) / 256, 256, 0, st>>>(p);
    yh_compact_rows<<<(m + wpb - 1) / wpb, wpb * 32, 0, st>>>(p);
    yh_compact_offsets<<<1, 32, 0, st>>>(p);
    yh_compact_scatter<<<(n + wpb - 1) / wpb, wpb * 32, 0, st>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_compact_targets launch");
}

}  // extern "C"
