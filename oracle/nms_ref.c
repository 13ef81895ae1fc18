/* Plain-C restatement of tf.image.non_max_suppression (TEST INFRASTRUCTURE).
 *
 * The reference calls it once per class from yolo_eval
 * (reference code/yolo3/model.py:474-480).  The op is TensorFlow's
 * NonMaxSuppressionV3 CPU kernel - third-party, un-vendored and version-unpinned
 * (code/README.md:12-14); its published algorithm is restated here:
 *   - candidates: score > score_threshold (strict);
 *   - visit order: score descending, ties -> lower box index first;
 *   - a candidate is selected unless IoU(candidate, s) > iou_threshold (strict)
 *     for an already selected s, scanning selected newest -> oldest;
 *   - stop at max_output_size;
 *   - IoU: corners normalised with min/max; 0 if either area <= 0;
 *     inter / (area_i + area_j - inter), all fp32, no FMA contraction
 *     (compile with -ffp-contract=off).
 * Parity is unpinned by the reference (oracle/__init__.py).
 */
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int idx; } cand_t;

static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

static float iou_tf(const float* bi, const float* bj) {
    const float ymin_i = fminf_(bi[0], bi[2]), xmin_i = fminf_(bi[1], bi[3]);
    const float ymax_i = fmaxf_(bi[0], bi[2]), xmax_i = fmaxf_(bi[1], bi[3]);
    const float ymin_j = fminf_(bj[0], bj[2]), xmin_j = fminf_(bj[1], bj[3]);
    const float ymax_j = fmaxf_(bj[0], bj[2]), xmax_j = fmaxf_(bj[1], bj[3]);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iy0 = fmaxf_(ymin_i, ymin_j), ix0 = fmaxf_(xmin_i, xmin_j);
    const float iy1 = fminf_(ymax_i, ymax_j), ix1 = fminf_(xmax_i, xmax_j);
    const float inter = fmaxf_(iy1 - iy0, 0.0f) * fmaxf_(ix1 - ix0, 0.0f);
    return inter / (area_i + area_j - inter);
}

/* boxes [n][4] (ymin,xmin,ymax,xmax); scores read at scores[i*score_stride].
 * Writes selected box indices (selection order) to out, returns their count. */
int nms_ref(const float* boxes, const float* scores, int n, int score_stride, int max_out,
            float iou_thr, float score_thr, int* out) {
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float s = scores[(size_t)i * (size_t)score_stride];
        if (s > score_thr) { c[m].score = s; c[m].idx = i; ++m; }
    }
    qsort(c, (size_t)m, sizeof(cand_t), cand_cmp);
    int k = 0;
    for (int t = 0; t < m && k < max_out; ++t) {
        const float* bi = boxes + 4 * (size_t)c[t].idx;
        int keep = 1;
        for (int j = k - 1; j >= 0; --j) {
            if (iou_tf(bi, boxes + 4 * (size_t)out[j]) > iou_thr) { keep = 0; break; }
        }
        if (keep) out[k++] = c[t].idx;
    }
    free(c);
    return k;
}
