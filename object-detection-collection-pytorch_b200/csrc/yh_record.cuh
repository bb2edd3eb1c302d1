// One ground-truth record against its cell, and the no-object arithmetic of one objectness logit: the per-record
// and per-logit arithmetic shared by the train head (yh_train.cu) and the fused step's image kernel (yh_nms.cu).
// Identical helpers -> identical bits: dL/dy does not depend on which kernel produced it.
#pragma once

#include <limits.h>

#include "yh_common.cuh"

constexpr int kClsRegs = 4;               // class logits per lane kept in registers (C <= 128)

struct WarpSums {
    float no, xy, wh, conf, nr, cls;
};

// order-preserving float <-> int map (for redux.sync max on floats; NaNs are not ordered)
__device__ __forceinline__ int yh_ordered(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float yh_unordered(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// No-object part of one objectness logit `t` in an image with `kn` boxes: the term kn * conf^2 and
// its gradient cno * kn * conf^2 * (1 - conf).  Approximate sigmoid (no decision depends on it) and
// explicitly rounded steps: the dense pass and the record warp must produce identical bits.
__device__ __forceinline__ float noobj_term(float t, float kn, float* conf_out) {
    // (two SFU operations: ex2.approx and rcp.approx handle +-inf, 0 and NaN the way the sigmoid needs them, and
    // the range fix-ups __expf / __fdividef wrap around them were a quarter of the dense pass' instructions)
    float e, conf;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(t, -1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(conf) : "f"(__fadd_rn(1.0f, e)));
    *conf_out = conf;
    return __fmul_rn(__fmul_rn(kn, conf), conf);
}
__device__ __forceinline__ float noobj_grad(float w, float conf, float cno) {
    return __fmul_rn(__fmul_rn(cno, w), __fsub_rn(1.0f, conf));
}

struct RecordRegs {
    int4 hd;    // img, cy, cx, cls
    float4 tt;  // stx, sty, tw, th
    float4 bb;  // x1, y1, x2, y2
};

// One ground-truth record against its cell; all 32 lanes cooperate and the work is laid out for
// LATENCY (records are the epilogue of a tile):
//   * lane l < 5A owns ONE activation (anchor l/5, channel l%5), so the 5A exp/sigmoid chains run
//     side by side; four shuffles hand every lane its anchor's box, the IoU is formed, and two
//     redux.sync steps pick the responsible anchor (max IoU, then lowest index: torch's first max);
//   * as soon as the anchor is known every lane issues its loads of the class logits and of the
//     dL/dy values it will update, so their latency hides behind the channel arithmetic;
//   * the five lanes of the responsible anchor each finish THEIR channel (x, y, w, h, conf:
//     target transform, squared error, gradient) in parallel; their squared errors accumulate in
//     per-lane registers by channel role;
//   * class softmax: lanes stride the classes, max through redux.sync on an order-preserving
//     integer image, then ONE butterfly for (sum e, sum e^2): with p = e / sum e,
//       sum_c (p_c - 1[c=t])^2 = S2 - 2 p_t + 1   and   sum_c (p_c - 1[c=t]) p_c = S2 - p_t,
//     S2 = sum p^2, so no third reduction is needed.
// `ycell` points at the cell's floats of y (in the shared-memory tile), `dcell` at the cell's
// floats of dy (global memory; holds the dense pass' values, visible after the CTA barrier);
// `kn` is the box count of the cell's image; `my_pw/my_ph` are the anchor multipliers of this
// lane's anchor (lane / 5).  Requires 5A <= 32.
// MODE 0: loss only; 1: add the gradient onto dy in global memory (`again`: an earlier record
// already updated this cell, so the row is read back; otherwise it holds the dense pass' values,
// which are known without a load: zero but the objectness channel); 2: leave the gradient in the
// shared-memory patch row `patch` (+ the dense objectness value of the row in *pdense), to be
// applied after the dense pass.
template <int MODE, class P>
__device__ __forceinline__ int process_record(const P& p, const int version, const int A, const int C,
                                              const RecordRegs& rr, int jj, const float* ycell,
                                              float* dcell, bool again, float* patch, float* pdense, float kn,
                                              int lane, float my_pw, float my_ph, WarpSums& s) {
    const YhGeom& g = p.g;
    const int bs = version == 2 ? 5 + C : 5;
    const int4 hd = rr.hd;
    const float4 tt = rr.tt;
    const float4 bb = rr.bb;

    const int a = lane / 5, q = lane - 5 * a;
    const bool mine = lane < 5 * A;
    float act = 0.f, t_raw = 0.f;
    if (mine) {
        const float t = ycell[a * bs + q];
        t_raw = t;
        const bool is_exp = version == 2 && (q == 2 || q == 3);
        const float e = expf(is_exp ? t : -t);
        act = is_exp ? e : __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
    }
    const int l0 = mine ? 5 * a : 0;
    const float bx_s = __shfl_sync(0xffffffffu, act, l0);
    const float by_s = __shfl_sync(0xffffffffu, act, l0 + 1);
    const float bw_a = __shfl_sync(0xffffffffu, act, l0 + 2);
    const float bh_a = __shfl_sync(0xffffffffu, act, l0 + 3);
    int key = INT_MIN;  // order-preserving image of this lane's IoU; NaN (as torch) ranks highest
    float iou = 0.f;
    if (mine) {
        const YhBox pb = yh_decode_box(bx_s, by_s, bw_a, bh_a, my_pw, my_ph, hd.z, hd.y, g.gw, g.gh);
        const YhBox gb{bb.x, bb.y, bb.z, bb.w};
        iou = yh_iou_xyxy(pb, gb);
        key = iou != iou ? INT_MAX : yh_ordered(iou);
    }
    const int best = __reduce_max_sync(0xffffffffu, key);
    const int r = __reduce_min_sync(0xffffffffu, (mine && key == best) ? a : 1 << 20);  // first max

    // loads that depend on r go out now: class logits (and, MODE 1, the dL/dy values to be updated)
    const int coff = version == 2 ? r * bs + 5 : 5 * A;
    const float* cl = ycell + coff;
    float* dcl = dcell + coff;
    const bool resp_lane = mine && a == r;
    float old_ch = 0.f;
    if (MODE == 1 && resp_lane) {
        if (again) {
            old_ch = dcell[r * bs + q];
        } else if (q == 4) {  // what the dense pass wrote for this logit (same helper, same bits)
            float cf_;
            const float w = noobj_term(t_raw, kn, &cf_);
            old_ch = noobj_grad(w, cf_, p.cno);
        }
    }
    float lg[kClsRegs], oldc[kClsRegs];
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        const int c = lane + 32 * k;
        lg[k] = c < C ? cl[c] : -INFINITY;
        oldc[k] = (MODE == 1 && again && c < C) ? dcl[c] : 0.f;
    }
    const float iou_r = __shfl_sync(0xffffffffu, iou, 5 * r);

    // the five lanes of the responsible anchor finish one channel each
    if (resp_lane) {
        float d, grad;
        if (q < 2) {            // x, y: (sigmoid(t) - target)^2, models/yolov2.py:1046-1050
            d = act - (q == 0 ? tt.x : tt.y);
            grad = p.cxy * d * act * (1.f - act);
            s.xy += d * d;
        } else if (q < 4) {     // w, h: (sqrt(act) - sqrt(target))^2, models/yolov2.py:946-947, 1063-1067
            const float t = q == 2 ? tt.z : tt.w;
            const float tgt = version == 2 ? __fsqrt_rn(__fdiv_rn(t, q == 2 ? my_pw : my_ph)) : __fsqrt_rn(t);
            const float qv = __fsqrt_rn(act);
            d = qv - tgt;
            grad = p.cwh * d * qv;
            if (version != 2) grad *= 1.f - act;  // v1: d sqrt(sigmoid(t)) / dt, models/yolov1.py:745-761
            s.wh += d * d;
        } else {                // objectness: (iou - conf)^2 and the no-object correction
            d = act - iou_r;
            grad = (p.cconf * d - p.cno * act) * act * (1.f - act);
            s.conf += d * d;
            s.nr += act * act;
            if (p.resp) p.resp[jj] = r;
            if (p.iou_resp) p.iou_resp[jj] = iou_r;
            if (MODE == 2) {  // what the dense pass writes for this logit
                float cf_;
                const float w = noobj_term(t_raw, kn, &cf_);
                *pdense = noobj_grad(w, cf_, p.cno);
            }
        }
        if (MODE == 1) dcell[r * bs + q] = __fadd_rn(old_ch, grad);  // (explicit add: MODE 1 and 2 must round alike)
        if (MODE == 2) patch[q] = grad;
    }

    // class term
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) mx = fmaxf(mx, lg[k]);
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) mx = fmaxf(mx, cl[c]);
    mx = yh_unordered(__reduce_max_sync(0xffffffffu, yh_ordered(mx)));
    float s1 = 0.f, s2 = 0.f, et = 0.f;  // sum e, sum e^2, e of the target class (owning lane only)
#pragma unroll
    for (int k = 0; k < kClsRegs; ++k) {
        const int c = lane + 32 * k;
        lg[k] = c < C ? expf(lg[k] - mx) : 0.f;
        s1 += lg[k];
        s2 += lg[k] * lg[k];
        if (c == hd.w) et = lg[k];
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
        const float e = expf(cl[c] - mx);
        s1 += e;
        s2 += e * e;
        if (c == hd.w) et = e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const bool has_t = hd.w >= 0 && hd.w < C;
    et = __shfl_sync(0xffffffffu, et, has_t ? (hd.w & 31) : 0);
    const float inv = __fdiv_rn(1.0f, s1);
    const float S2 = s2 * inv * inv;
    const float pt = has_t ? et * inv : 0.f;
    const float dot = S2 - pt;
    if (lane == 0) s.cls += S2 - 2.f * pt + (has_t ? 1.f : 0.f);
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < kClsRegs; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float pc = lg[k] * inv;
                dcl[c] = __fadd_rn(oldc[k], p.ccls * pc * (pc - (c == hd.w ? 1.f : 0.f) - dot));
            }
        }
        for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
            const float pc = expf(cl[c] - mx) * inv;
            dcl[c] = __fadd_rn(again ? dcl[c] : 0.f, p.ccls * pc * (pc - (c == hd.w ? 1.f : 0.f) - dot));
        }
    }
    if (MODE == 2 && lane < C) {  // 5 + C <= kPatchFloats: one class per lane
        const float pc = lg[0] * inv;
        patch[5 + lane] = p.ccls * pc * (pc - (lane == hd.w ? 1.f : 0.f) - dot);
    }
    __syncwarp();
    return r;
}

