// Device-side target builder: pixel boxes -> compact ground-truth records + per-image CSR offsets.
//
// Replaces the per-box arithmetic of the reference's collate_fn (models/yolov2.py:1440-1512,
// models/yolov1.py:1238-1312), which runs in numpy float64 on the host and then scatters every
// box into its own dense [S_h,S_w,.] grids (2.2 GB at BASELINE config 5).  Here one thread per
// box computes the same 12 scalars in float64, in the reference's operation order (IEEE divisions,
// no fused multiply-adds), casts once to float32 like `torch.tensor(...).float()`, and writes the
// 48-byte record the train head consumes; the boxes arrive grouped by image (collate_fn appends
// them image by image), so the CSR offsets are one binary search per image.  One launch.
#include "yh_common.cuh"

namespace {

struct BuildParams {
    const double* boxes;     // [m,4] x1,y1,x2,y2 pixels
    const int32_t* labels;   // [m]
    const int32_t* img;      // [m] position of the owning image in the batch, non-decreasing
    int m, n, version, s_h, s_w;
    double gh, gw;           // pixels per grid cell: H / S_h, W / S_w (Python floats)
    YhGt* gt;
    int32_t* gt_off;         // [n+1]
    int32_t* status;         // [2]: boxes out of order / with an image index or cell out of range
};

__global__ void yh_build_targets_zero(BuildParams p) {
    if (threadIdx.x < 2) p.status[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256) yh_build_targets_kernel(BuildParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p.m) {
        const double x1 = p.boxes[4 * j], y1 = p.boxes[4 * j + 1], x2 = p.boxes[4 * j + 2], y2 = p.boxes[4 * j + 3];
        // models/yolov2.py:1466-1482, one rounding per operation
        const double x1n = __ddiv_rn(x1, p.gw), y1n = __ddiv_rn(y1, p.gh);
        const double x2n = __ddiv_rn(x2, p.gw), y2n = __ddiv_rn(y2, p.gh);
        const double bx = __ddiv_rn(__dadd_rn(x1n, x2n), 2.0), by = __ddiv_rn(__dadd_rn(y1n, y2n), 2.0);
        const double bw = __dsub_rn(x2n, x1n), bh = __dsub_rn(y2n, y1n);
        const int cx = (int)bx, cy = (int)by;  // int(): truncation
        YhGt r;
        r.img = p.img[j];
        r.cy = cy;
        r.cx = cx;
        r.cls = p.labels[j];
        r.stx = (float)__dsub_rn(bx, (double)cx);
        r.sty = (float)__dsub_rn(by, (double)cy);
        if (p.version == 1) {  // models/yolov1.py:1281-1282
            r.tw = (float)__ddiv_rn(bw, (double)p.s_w);
            r.th = (float)__ddiv_rn(bh, (double)p.s_h);
        } else {
            r.tw = (float)bw;
            r.th = (float)bh;
        }
        r.x1 = (float)x1; r.y1 = (float)y1; r.x2 = (float)x2; r.y2 = (float)y2;
        p.gt[j] = r;
        if (j > 0 && p.img[j - 1] > r.img) atomicAdd(p.status, 1);
        if (r.img < 0 || r.img >= p.n || cx < 0 || cx >= p.s_w || cy < 0 || cy >= p.s_h) atomicAdd(p.status + 1, 1);
    }
    if (j <= p.n) {  // gt_off[j] = first box whose image index is >= j
        int lo = 0, hi = p.m;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p.img[mid] < j) lo = mid + 1; else hi = mid;
        }
        p.gt_off[j] = lo;
    }
}

}  // namespace

extern "C" {

int yh_build_targets(const double* boxes_xyxy, const int32_t* labels, const int32_t* img_index, int m, int n,
                     int version, int s_h, int s_w, double img_h, double img_w, YhGt* gt_out,
                     int32_t* gt_off_out, int32_t* status, void* stream) {
    YH_REQUIRE(version == 1 || version == 2, YH_ERR_INVALID, "version must be 1 or 2");
    YH_REQUIRE(m > 0, YH_ERR_EMPTY, "no ground-truth boxes (m=%d)", m);
    YH_REQUIRE(n > 0 && s_h > 0 && s_w > 0 && img_h > 0 && img_w > 0, YH_ERR_INVALID, "bad sizes");
    YH_REQUIRE(boxes_xyxy && labels && img_index && gt_out && gt_off_out && status, YH_ERR_INVALID, "null pointer argument");
    YH_REQUIRE(((uintptr_t)gt_out & 15) == 0 && ((uintptr_t)boxes_xyxy & 7) == 0, YH_ERR_INVALID, "misaligned pointer");
    BuildParams p;
    p.boxes = boxes_xyxy; p.labels = labels; p.img = img_index;
    p.m = m; p.n = n; p.version = version; p.s_h = s_h; p.s_w = s_w;
    p.gh = img_h / (double)s_h;
    p.gw = img_w / (double)s_w;
    p.gt = gt_out; p.gt_off = gt_off_out; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    yh_build_targets_zero<<<1, 32, 0, st>>>(p);
    const int work = m > n + 1 ? m : n + 1;
    yh_build_targets_kernel<<<(work + 255) / 256, 256, 0, st>>>(p);
    return yh_check_cuda(cudaGetLastError(), "yh_build_targets launch");
}

}  // extern "C"
