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


// ---- arithmetic shared by the one-record-per-warp form (process_record) and the four-records-per-warp form
// (process_records4).  Every step that enters dL/dy is an explicitly rounded operation (nothing is left to the
// compiler's FMA contraction), so both forms -- and every kernel that uses them -- produce the same bits.

// activation of one box logit: exp(t) for the v2 sizes, sigmoid(t) otherwise
__device__ __forceinline__ float yh_box_act(float t, bool is_exp) {
    const float e = expf(is_exp ? t : -t);
    return is_exp ? e : __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// Channel q (x, y, w, h, objectness) of the responsible predictor: the squared error goes into the lane's sums,
// the gradient d(loss)/d(logit) is returned.  `act` is the channel's activation, `pw`/`ph` the anchor multipliers
// of the responsible predictor, `iou_r` its IoU with the record's box.
template <class P>
__device__ __forceinline__ float yh_channel(const P& p, const int version, const int q, const float act,
                                            const float4& tt, const float pw, const float ph, const float iou_r,
                                            WarpSums& s) {
    const float om = __fsub_rn(1.0f, act);
    if (q < 2) {            // x, y: (sigmoid(t) - target)^2, models/yolov2.py:1046-1050
        const float d = __fsub_rn(act, q == 0 ? tt.x : tt.y);
        s.xy = fmaf(d, d, s.xy);
        return __fmul_rn(__fmul_rn(__fmul_rn(p.cxy, d), act), om);
    }
    if (q < 4) {            // w, h: (sqrt(act) - sqrt(target))^2, models/yolov2.py:946-947, 1063-1067
        const float t = q == 2 ? tt.z : tt.w;
        const float tgt = version == 2 ? __fsqrt_rn(__fdiv_rn(t, q == 2 ? pw : ph)) : __fsqrt_rn(t);
        const float qv = __fsqrt_rn(act);
        const float d = __fsub_rn(qv, tgt);
        s.wh = fmaf(d, d, s.wh);
        const float gr = __fmul_rn(__fmul_rn(p.cwh, d), qv);
        return version == 2 ? gr : __fmul_rn(gr, om);  // v1: d sqrt(sigmoid(t)) / dt, models/yolov1.py:745-761
    }
    // objectness: (iou - conf)^2 and the no-object correction
    const float d = __fsub_rn(act, iou_r);
    s.conf = fmaf(d, d, s.conf);
    s.nr = fmaf(act, act, s.nr);
    return __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(p.cconf, d), __fmul_rn(p.cno, act)), act), om);
}

// softmax statistics -> what the class gradients need: with p = e / s1,  S2 = sum p^2,  pt = p of the target class
struct YhClsStat {
    float inv, S2, pt, dot;
};
__device__ __forceinline__ YhClsStat yh_cls_stat(float s1, float s2, float et, bool has_t) {
    YhClsStat k;
    k.inv = __fdiv_rn(1.0f, s1);
    k.S2 = __fmul_rn(__fmul_rn(s2, k.inv), k.inv);
    k.pt = has_t ? __fmul_rn(et, k.inv) : 0.f;
    k.dot = __fsub_rn(k.S2, k.pt);
    return k;
}
// the class term's contribution  sum_c (p_c - 1[c=t])^2 = S2 - 2 pt + 1
__device__ __forceinline__ float yh_cls_term(const YhClsStat& k, bool has_t) {
    return __fadd_rn(__fsub_rn(k.S2, __fmul_rn(2.f, k.pt)), has_t ? 1.f : 0.f);
}
// d(class term)/d(logit c) for e = exp(logit_c - max)
__device__ __forceinline__ float yh_cls_grad(float ccls, float e, const YhClsStat& k, bool is_target) {
    const float pc = __fmul_rn(e, k.inv);
    return __fmul_rn(__fmul_rn(ccls, pc), __fsub_rn(__fsub_rn(pc, is_target ? 1.f : 0.f), k.dot));
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
        act = yh_box_act(t, version == 2 && (q == 2 || q == 3));
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
        const float grad = yh_channel(p, version, q, act, tt, my_pw, my_ph, iou_r, s);
        if (q == 4) {
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
        s1 = __fadd_rn(s1, lg[k]);
        s2 = fmaf(lg[k], lg[k], s2);
        if (c == hd.w) et = lg[k];
    }
    for (int c = lane + 32 * kClsRegs; c < C; c += 32) {
        const float e = expf(cl[c] - mx);
        s1 = __fadd_rn(s1, e);
        s2 = fmaf(e, e, s2);
        if (c == hd.w) et = e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 = __fadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, o));
        s2 = __fadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, o));
    }
    const bool has_t = hd.w >= 0 && hd.w < C;
    et = __shfl_sync(0xffffffffu, et, has_t ? (hd.w & 31) : 0);
    const YhClsStat ks = yh_cls_stat(s1, s2, et, has_t);
    if (lane == 0) s.cls += yh_cls_term(ks, has_t);
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < kClsRegs; ++k) {
            const int c = lane + 32 * k;
            if (c < C) dcl[c] = __fadd_rn(oldc[k], yh_cls_grad(p.ccls, lg[k], ks, c == hd.w));
        }
        for (int c = lane + 32 * kClsRegs; c < C; c += 32)
            dcl[c] = __fadd_rn(again ? dcl[c] : 0.f, yh_cls_grad(p.ccls, expf(cl[c] - mx), ks, c == hd.w));
    }
    if (MODE == 2 && lane < C) patch[5 + lane] = yh_cls_grad(p.ccls, lg[0], ks, lane == hd.w);  // 5 + C <= kPatchFloats: one class per lane
    __syncwarp();
    return r;
}

// Four records per warp, eight lanes each -- for inputs dense with ground truth (BASELINE config 5: 50-100 boxes per
// image), where the records' instruction count, not their latency, is what the kernel waits for.  The one-warp form
// spends ~440 warp instructions per record with 25 of 32 lanes busy in its widest phase and one or five in most
// others; here a group of eight lanes does one record:
//   * lane sub < A of a group owns one ANCHOR: its four box activations (independent chains, issued back to back),
//     the decoded box, the IoU; two redux.sync steps over the group's eight lanes pick the responsible anchor;
//   * the activations of that anchor are handed to lanes 0..3 of the group by four shuffles (one instruction each for
//     all four records), lane 4 activates the objectness logit: five lanes finish one channel each, as before;
//   * class softmax: lane sub covers classes sub, sub + 8, ...; the partial sums are kept per VIRTUAL lane of the
//     one-warp form (class c belongs to virtual lane c mod 32) and folded in that form's butterfly order
//     (xor 16 and xor 8 inside the lane, xor 4, 2, 1 by shuffles), so s1 and s2 have the same bits.
// Same helpers, same roundings, same addition trees as process_record: dL/dy, resp and iou_resp are bit-identical
// whichever form processed a record (tests/test_gpu_parity.py compares them).  MODE 0 (loss only) and 1 (onto dy).
// `on`: this lane's group has a record (idle groups take part in the shuffles); `rr`, `jj`, `ycell`, `dcell`,
// `again`, `kn` are the group's.  The records of one call must lie in different cells (the caller splits batches).
constexpr int kCls4Regs = 4;  // class logits per lane kept in registers (C <= 32; more: recomputed)
template <int MODE, class P>
__device__ __forceinline__ void process_records4(const P& p, const int version, const int A, const int C,
                                                 const bool on, const RecordRegs& rr, const int jj,
                                                 const float* ycell, float* dcell, const bool again, const float kn,
                                                 const int lane, WarpSums& s) {
    static_assert(MODE == 0 || MODE == 1, "patches are produced by the one-warp form");
    const YhGeom& g = p.g;
    const int bs = version == 2 ? 5 + C : 5;
    const int sub = lane & 7, l0 = lane & 24;
    const unsigned gmask = 0xffu << l0;
    const int4 hd = rr.hd;
    const float4 tt = rr.tt;
    const float4 bb = rr.bb;

    const bool mine = on && sub < A;
    float ax = 0.f, ay = 0.f, aw = 0.f, ah = 0.f, iou = 0.f;
    int key = INT_MIN;
    if (mine) {
        const float* b = ycell + sub * bs;
        const float tx = b[0], ty = b[1], tw = b[2], th = b[3];
        ax = yh_box_act(tx, false);
        ay = yh_box_act(ty, false);
        aw = yh_box_act(tw, version == 2);
        ah = yh_box_act(th, version == 2);
        const YhBox pb = yh_decode_box(ax, ay, aw, ah, g.pw[sub], g.ph[sub], hd.z, hd.y, g.gw, g.gh);
        const YhBox gb{bb.x, bb.y, bb.z, bb.w};
        iou = yh_iou_xyxy(pb, gb);
        key = iou != iou ? INT_MAX : yh_ordered(iou);
    }
    const int best = __reduce_max_sync(gmask, key);
    int r = __reduce_min_sync(gmask, (mine && key == best) ? sub : 1 << 20);  // first max
    if (!on) r = 0;

    // loads that depend on r go out now
    const int coff = version == 2 ? r * bs + 5 : 5 * A;
    const float* cl = ycell + coff;
    float* dcl = dcell + coff;
    const int q = sub;
    const bool ch_lane = on && q < 5;
    float t4 = 0.f, old_ch = 0.f;
    if (on && q == 4) t4 = ycell[r * bs + 4];
    if (MODE == 1 && ch_lane && again) old_ch = dcell[r * bs + q];
    float e[kCls4Regs], oldc[kCls4Regs];
#pragma unroll
    for (int k = 0; k < kCls4Regs; ++k) {
        const int c = sub + 8 * k;
        e[k] = (on && c < C) ? cl[c] : -INFINITY;
        oldc[k] = (MODE == 1 && again && on && c < C) ? dcl[c] : 0.f;
    }
    const int src = l0 + r;
    const float sx = __shfl_sync(0xffffffffu, ax, src), sy = __shfl_sync(0xffffffffu, ay, src);
    const float sw = __shfl_sync(0xffffffffu, aw, src), sh = __shfl_sync(0xffffffffu, ah, src);
    const float iou_r = __shfl_sync(0xffffffffu, iou, src);

    if (ch_lane) {
        float act = sh;
        act = q == 2 ? sw : act;
        act = q == 1 ? sy : act;
        act = q == 0 ? sx : act;
        if (q == 4) {
            act = yh_box_act(t4, false);
            if (MODE == 1 && !again) {  // what the dense pass wrote for this logit (same helper, same bits)
                float cf_;
                const float w = noobj_term(t4, kn, &cf_);
                old_ch = noobj_grad(w, cf_, p.cno);
            }
        }
        const float grad = yh_channel(p, version, q, act, tt, g.pw[r], g.ph[r], iou_r, s);
        if (q == 4) {
            if (p.resp) p.resp[jj] = r;
            if (p.iou_resp) p.iou_resp[jj] = iou_r;
        }
        if (MODE == 1) dcell[r * bs + q] = __fadd_rn(old_ch, grad);
    }

    // class term
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kCls4Regs; ++k) mx = fmaxf(mx, e[k]);
    if (on)
        for (int c = sub + 8 * kCls4Regs; c < C; c += 8) mx = fmaxf(mx, cl[c]);
    mx = yh_unordered(__reduce_max_sync(gmask, yh_ordered(mx)));
    // partial sums per virtual lane sub + 8 m of the one-warp form (class c -> m = (c / 8) & 3), in class order
    float v1[4], v2[4];
#pragma unroll
    for (int k = 0; k < kCls4Regs; ++k) {
        const int c = sub + 8 * k;
        e[k] = (on && c < C) ? expf(e[k] - mx) : 0.f;
        v1[k] = e[k];
        v2[k] = __fmul_rn(e[k], e[k]);
    }
    if (on && C > 8 * kCls4Regs) {
        for (int c0 = 8 * kCls4Regs; c0 < C; c0 += 32) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int c = c0 + 8 * m + sub;
                if (c < C) {
                    const float ee = expf(cl[c] - mx);
                    v1[m] = __fadd_rn(v1[m], ee);
                    v2[m] = fmaf(ee, ee, v2[m]);
                }
            }
        }
    }
    float s1 = __fadd_rn(__fadd_rn(v1[0], v1[2]), __fadd_rn(v1[1], v1[3]));  // xor 16, then xor 8
    float s2 = __fadd_rn(__fadd_rn(v2[0], v2[2]), __fadd_rn(v2[1], v2[3]));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        s1 = __fadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, o));
        s2 = __fadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, o));
    }
    const bool has_t = hd.w >= 0 && hd.w < C;
    if (on) {
        const float et = has_t ? expf(cl[hd.w] - mx) : 0.f;
        const YhClsStat ks = yh_cls_stat(s1, s2, et, has_t);
        if (sub == 0) s.cls += yh_cls_term(ks, has_t);
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < kCls4Regs; ++k) {
                const int c = sub + 8 * k;
                if (c < C) dcl[c] = __fadd_rn(oldc[k], yh_cls_grad(p.ccls, e[k], ks, c == hd.w));
            }
            for (int c = sub + 8 * kCls4Regs; c < C; c += 8)
                dcl[c] = __fadd_rn(again ? dcl[c] : 0.f, yh_cls_grad(p.ccls, expf(cl[c] - mx), ks, c == hd.w));
        }
    }
    __syncwarp();
}

// The next batch of up to four records for process_records4 out of the ballot `bal` of a warp's pending records
// (bit b: the record held by lane b; `lc` the lane's tile-/image-local cell): group g takes the g-th lowest bit.
// A record whose cell equals that of an earlier record of the batch ends the batch in front of it (records of one
// cell accumulate in CSR order, one after the other).  Returns the lane b the group's record comes from (-1: the
// group idles) and removes the batch's bits from `bal`.
__device__ __forceinline__ int yh_batch4(unsigned& bal, const int lc, const int lane) {
    const int grp = lane >> 3;
    int b = -1;
    unsigned rest = bal;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int bk = rest ? __ffs(rest) - 1 : -1;
        if (k == grp) b = bk;
        rest &= rest - 1u;  // (0 stays 0)
    }
    const int mylc = __shfl_sync(0xffffffffu, lc, b >= 0 ? b : 0);
    // first group whose cell repeats an earlier group's: it and the groups behind it wait for the next batch
    bool dup = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int lk = __shfl_sync(0xffffffffu, mylc, 8 * k);
        const int bk = __shfl_sync(0xffffffffu, b, 8 * k);
        dup = dup || (k < grp && b >= 0 && bk >= 0 && lk == mylc);
    }
    const unsigned dm = __ballot_sync(0xffffffffu, dup);
    const int cut = dm ? (__ffs(dm) - 1) >> 3 : 4;  // groups < cut run now
    if (grp >= cut) b = -1;
    // remove the bits of the groups that run
    unsigned taken = 0u;
    unsigned rest2 = bal;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < cut && rest2) taken |= rest2 & (0u - rest2);
        rest2 &= rest2 - 1u;
    }
    bal &= ~taken;
    return b;
}
