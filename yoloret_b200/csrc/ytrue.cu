// y_true encoder on the device: reference preprocess_true_boxes (code/yolo3/utils.py:298-376), which the
// reference runs as a numpy py_function inside its tf.data pipeline (code/yolo3/data.py:84-121).
//
//   boxes [B][T][5] = (xmin, ymin, xmax, ymax, class) in input pixels, zero-width rows are padding
//   ->  per scale l a dense [B][gh][gw][3][5+C] tensor: (cx, cy, w, h) normalised, objectness, one-hot class,
//       written at the cell holding the floor-divided box centre for the anchor (of all 9) whose SHAPE has the best
//       IoU with the box.
// The reference's loop is sequential per image: a later box overwrites an earlier one in the same slot but leaves
// the earlier class bit set, and its counter over the VALID boxes indexes the UNFILTERED rows (utils.py:357-368).
// Both are reproduced: one thread walks one image's boxes in order; images are independent.  The tensors are
// cleared by one cudaMemsetAsync each (that is the HBM traffic of this op; the scatter itself is a few KB).
#include "yr_common.cuh"

namespace yr {

struct YTrueArgs {
    const float* boxes;
    float* y[3];
    int gh[3], gw[3];
    float anchors[18];
    int B, T, in_h, in_w, C, num_scales;
};

__global__ void __launch_bounds__(128)
encode_true_boxes_kernel(const YTrueArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const float* tb = a.boxes + (size_t)b * a.T * 5;
    const int E = 5 + a.C;
    int k = 0;  // counter over the valid boxes; also the row the reference reads position and class from
    for (int t = 0; t < a.T; ++t) {
        const float w = __fsub_rn(tb[t * 5 + 2], tb[t * 5 + 0]);
        if (!(w > 0.0f)) continue;                                  // valid_mask = boxes_wh[..., 0] > 0
        const float h = __fsub_rn(tb[t * 5 + 3], tb[t * 5 + 1]);
        int best = 0;
        float best_iou = -1.0f;
        for (int n = 0; n < 9; ++n) {                               // IoU of centred shapes, fp32, first maximum
            const float aw = a.anchors[2 * n], ah = a.anchors[2 * n + 1];
            const float inter = __fmul_rn(fminf(w, aw), fminf(h, ah));
            const float uni = __fsub_rn(__fadd_rn(__fmul_rn(w, h), __fmul_rn(aw, ah)), inter);
            const float iou = __fdiv_rn(inter, uni);
            if (iou > best_iou) { best_iou = iou; best = n; }
        }
        const float* row = tb + k * 5;                              // true_boxes[t] with t = index among valid boxes
        ++k;
        // (xmin + xmax) // 2 in float32, then / input size in float64, stored as float32 (numpy promotion rules)
        const float cx = floorf(__fmul_rn(__fadd_rn(row[0], row[2]), 0.5f));
        const float cy = floorf(__fmul_rn(__fadd_rn(row[1], row[3]), 0.5f));
        const float rw = __fsub_rn(row[2], row[0]), rh = __fsub_rn(row[3], row[1]);
        const float rel[4] = {(float)((double)cx / (double)a.in_w), (float)((double)cy / (double)a.in_h),
                              (float)((double)rw / (double)a.in_w), (float)((double)rh / (double)a.in_h)};
        const int l = 2 - best / 3;                                 // anchor_mask = [[6,7,8],[3,4,5],[0,1,2]]
        const int ls = l - (3 - a.num_scales);                      // anchor_mask[-num_scales:]
        if (ls < 0) continue;
        const int i = (int)floor((double)rel[0] * (double)a.gw[ls]);
        const int j = (int)floor((double)rel[1] * (double)a.gh[ls]);
        const int c = (int)row[4];
        if (i < 0 || i >= a.gw[ls] || j < 0 || j >= a.gh[ls] || c < 0 || c >= a.C) continue;  // numpy would raise
        float* dst = a.y[ls] + ((((size_t)b * a.gh[ls] + j) * a.gw[ls] + i) * 3 + (best % 3)) * E;
        dst[0] = rel[0]; dst[1] = rel[1]; dst[2] = rel[2]; dst[3] = rel[3];
        dst[4] = 1.0f;
        dst[5 + c] = 1.0f;
    }
}

}  // namespace yr

using namespace yr;

extern "C" int yr_encode_true_boxes(const float* boxes, int B, int T, const float* anchors_host, int in_h, int in_w,
                                    int num_classes, int num_scales, float* const* y_true, void* stream) {
    YR_CHECK_ARG((boxes || T == 0) && anchors_host && y_true, "encode_true_boxes: null pointer");
    YR_CHECK_ARG(B > 0 && T >= 0 && num_classes > 0 && num_scales >= 1 && num_scales <= 3 && in_h > 0 && in_w > 0,
                 "encode_true_boxes: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    YTrueArgs a;
    a.boxes = boxes;
    a.B = B;
    a.T = T;
    a.in_h = in_h;
    a.in_w = in_w;
    a.C = num_classes;
    a.num_scales = num_scales;
    for (int i = 0; i < 18; ++i) a.anchors[i] = anchors_host[i];
    const int steps[3] = {32, 16, 8};
    for (int l = 0; l < 3; ++l) {
        a.y[l] = nullptr;
        a.gh[l] = a.gw[l] = 0;
    }
    for (int l = 0; l < num_scales; ++l) {
        YR_CHECK_ARG(y_true[l] != nullptr, "encode_true_boxes: null y_true[%d]", l);
        a.y[l] = y_true[l];
        a.gh[l] = (int)nearbyint((double)in_h / steps[l]);   // np.round: half to even
        a.gw[l] = (int)nearbyint((double)in_w / steps[l]);
        const size_t bytes = (size_t)B * a.gh[l] * a.gw[l] * 3 * (5 + num_classes) * sizeof(float);
        if (cudaMemsetAsync(a.y[l], 0, bytes, s) != cudaSuccess) {
            set_error("encode_true_boxes: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
    }
    if (T > 0) {
        encode_true_boxes_kernel<<<cdiv(B, 128), 128, 0, s>>>(a);
        YR_CHECK_LAUNCH("encode_true_boxes");
    }
    return YR_OK;
}
