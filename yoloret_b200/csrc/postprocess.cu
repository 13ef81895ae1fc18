// yolo_eval on the GPU: fused box decode + score filter, class-wise greedy NMS, packing.
//
// Reference: yolo_head / yolo_correct_boxes / yolo_boxes_and_scores / yolo_eval,
// code/yolo3/model.py:344-491, with tf.image.non_max_suppression (NonMaxSuppressionV3)
// semantics restated in oracle/nms_ref.c.  This translation unit is compiled with
// -fmad=false: every multiply/add is a separately rounded fp32 op, in the reference's
// order, so NMS decisions are bit-identical to the CPU oracle on identical inputs.
#include "yr_common.cuh"
#include <math.h>

namespace yr {

constexpr int DEC_BOXES = 64;    // boxes per CTA
constexpr int DEC_THREADS = 256; // 8 warps x 8 boxes

struct DecodeArgs {
    const float* feats[3];
    int gh[3], gw[3], ld[3], box_off[4];
    float anchors[3][3][2];
    int num_scales, C, total_boxes, cand_cap;
    float in_h, in_w, thr;
};

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(DEC_THREADS)
decode_filter_kernel(DecodeArgs a, const float* __restrict__ image_shapes, float* __restrict__ boxes,
                     float* __restrict__ cand_score, int32_t* __restrict__ cand_index, int32_t* __restrict__ cand_count) {
    extern __shared__ float s_score[];  // [DEC_BOXES][C]
    const int b = blockIdx.y;
    const int box0 = blockIdx.x * DEC_BOXES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.C;
    const float img_h = __ldg(image_shapes + 2 * b), img_w = __ldg(image_shapes + 2 * b + 1);
    // yolo_correct_boxes constants (model.py:379-385)
    const float max_shape = fmaxf(img_h, img_w);
    const float boxed_h = a.in_h * (img_h / max_shape), boxed_w = a.in_w * (img_w / max_shape);
    const float off_h = (a.in_h - boxed_h) / 2.0f, off_w = (a.in_w - boxed_w) / 2.0f;
    const float scale_h = img_h / boxed_h, scale_w = img_w / boxed_w;

    for (int i = 0; i < DEC_BOXES / 8; ++i) {
        const int lb = warp * (DEC_BOXES / 8) + i;
        const int box = box0 + lb;
        if (box >= a.total_boxes) {
            for (int c = lane; c < C; c += 32) s_score[lb * C + c] = -INFINITY;
            continue;
        }
        int s = 0;
        if (a.num_scales > 1 && box >= a.box_off[1]) s = 1;
        if (a.num_scales > 2 && box >= a.box_off[2]) s = 2;
        const int local = box - a.box_off[s];
        const int anchor = local % 3, cell = local / 3;
        const int gx = cell % a.gw[s], gy = cell / a.gw[s];
        const float* f = a.feats[s] + ((size_t)b * a.gh[s] * a.gw[s] + cell) * a.ld[s] + anchor * (C + 5);
        const float v0 = lane < C + 5 ? __ldg(f + lane) : 0.0f;
        const float tx = __shfl_sync(0xffffffffu, v0, 0), ty = __shfl_sync(0xffffffffu, v0, 1);
        const float tw = __shfl_sync(0xffffffffu, v0, 2), th = __shfl_sync(0xffffffffu, v0, 3);
        const float conf = sigmoid_exact(__shfl_sync(0xffffffffu, v0, 4));
        if (lane == 0) {
            // yolo_head (model.py:363-366)
            const float bx = (sigmoid_exact(tx) + (float)gx) / (float)a.gw[s];
            const float by = (sigmoid_exact(ty) + (float)gy) / (float)a.gh[s];
            const float bw = expf(tw) * a.anchors[s][anchor][0] / a.in_w;
            const float bh = expf(th) * a.anchors[s][anchor][1] / a.in_h;
            // yolo_correct_boxes (model.py:386-398)
            const float y = (by * a.in_h - off_h) * scale_h;
            const float x = (bx * a.in_w - off_w) * scale_w;
            const float hh = bh * (a.in_h * scale_h);
            const float ww = bw * (a.in_w * scale_w);
            float4 o;
            o.x = fminf(fmaxf(y - hh / 2.0f, 0.0f), img_h);
            o.y = fminf(fmaxf(x - ww / 2.0f, 0.0f), img_w);
            o.z = fminf(fmaxf(y + hh / 2.0f, 0.0f), img_h);
            o.w = fminf(fmaxf(x + ww / 2.0f, 0.0f), img_w);
            *reinterpret_cast<float4*>(boxes + ((size_t)b * a.total_boxes + box) * 4) = o;
        }
        // box_scores = box_confidence * box_class_probs (model.py:426)
        for (int e = lane; e < C + 5; e += 32) {
            const float v = e < 32 ? v0 : __ldg(f + e);
            if (e >= 5) s_score[lb * C + (e - 5)] = conf * sigmoid_exact(v);
        }
    }
    __syncthreads();
    // per class: ordered compaction of this CTA's 64 boxes, one atomic per (CTA, class)
    for (int c = threadIdx.x; c < C; c += DEC_THREADS) {
        int n = 0;
        for (int i = 0; i < DEC_BOXES; ++i) n += (s_score[i * C + c] > a.thr) ? 1 : 0;
        if (n == 0) continue;
        const size_t list = (size_t)b * C + c;
        int pos = atomicAdd(cand_count + list, n);
        for (int i = 0; i < DEC_BOXES; ++i) {
            const float sc = s_score[i * C + c];
            if (sc > a.thr) {
                if (pos < a.cand_cap) {
                    cand_score[list * a.cand_cap + pos] = sc;
                    cand_index[list * a.cand_cap + pos] = box0 + i;
                }
                ++pos;
            }
        }
    }
}

// Sparse variant for score_threshold > 0 (the YOLO default is 0.2): score = conf * cls <= conf, so an
// anchor whose objectness fails the threshold cannot produce a candidate and its C class logits are
// never touched.  One warp per grid cell: the cell's A*(5+C) logits are one contiguous, 16-byte
// aligned run, read with 128-bit loads; candidates are appended with one atomic each (their order
// within a list is irrelevant: NMS orders by (score, index)).  Same arithmetic, op for op, as
// decode_filter_kernel - only the work distribution differs.
__global__ void __launch_bounds__(256)
decode_sparse_kernel(DecodeArgs a, const float* __restrict__ image_shapes, float* __restrict__ boxes,
                     float* __restrict__ cand_score, int32_t* __restrict__ cand_index, int32_t* __restrict__ cand_count,
                     int total_cells) {
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp_global >= total_cells) return;
    const int C = a.C, E = C + 5;
    int s = 0, cell = warp_global, cell_off = 0;
    const int n0 = a.gh[0] * a.gw[0], n1 = a.num_scales > 1 ? a.gh[1] * a.gw[1] : 0;
    if (a.num_scales > 1 && cell >= n0) { s = 1; cell -= n0; cell_off = n0; }
    if (a.num_scales > 2 && s == 1 && cell >= n1) { s = 2; cell -= n1; cell_off = n0 + n1; }
    (void)cell_off;
    const float* f = a.feats[s] + ((size_t)b * a.gh[s] * a.gw[s] + cell) * a.ld[s];
    // objectness of the 3 anchors
    float conf = 0.0f;
    if (lane < 3) conf = sigmoid_exact(__ldg(f + lane * E + 4));
    const unsigned mask = __ballot_sync(0xffffffffu, lane < 3 && conf > a.thr);
    if (mask == 0u) return;
    const float c0 = __shfl_sync(0xffffffffu, conf, 0), c1 = __shfl_sync(0xffffffffu, conf, 1);
    const float c2 = __shfl_sync(0xffffffffu, conf, 2);
    const int box_base = a.box_off[s] + cell * 3;
    if (lane < 3 && ((mask >> lane) & 1u)) {
        const float img_h = __ldg(image_shapes + 2 * b), img_w = __ldg(image_shapes + 2 * b + 1);
        const float max_shape = fmaxf(img_h, img_w);
        const float boxed_h = a.in_h * (img_h / max_shape), boxed_w = a.in_w * (img_w / max_shape);
        const float off_h = (a.in_h - boxed_h) / 2.0f, off_w = (a.in_w - boxed_w) / 2.0f;
        const float scale_h = img_h / boxed_h, scale_w = img_w / boxed_w;
        const int gx = cell % a.gw[s], gy = cell / a.gw[s];
        const float* t = f + lane * E;
        const float tx = __ldg(t), ty = __ldg(t + 1), tw = __ldg(t + 2), th = __ldg(t + 3);
        const float bx = (sigmoid_exact(tx) + (float)gx) / (float)a.gw[s];
        const float by = (sigmoid_exact(ty) + (float)gy) / (float)a.gh[s];
        const float bw = expf(tw) * a.anchors[s][lane][0] / a.in_w;
        const float bh = expf(th) * a.anchors[s][lane][1] / a.in_h;
        const float y = (by * a.in_h - off_h) * scale_h;
        const float x = (bx * a.in_w - off_w) * scale_w;
        const float hh = bh * (a.in_h * scale_h);
        const float ww = bw * (a.in_w * scale_w);
        float4 o;
        o.x = fminf(fmaxf(y - hh / 2.0f, 0.0f), img_h);
        o.y = fminf(fmaxf(x - ww / 2.0f, 0.0f), img_w);
        o.z = fminf(fmaxf(y + hh / 2.0f, 0.0f), img_h);
        o.w = fminf(fmaxf(x + ww / 2.0f, 0.0f), img_w);
        *reinterpret_cast<float4*>(boxes + ((size_t)b * a.total_boxes + box_base + lane) * 4) = o;
    }
    // class scores of the surviving anchors
    const int valid = 3 * E;
    for (int e0 = lane * 4; e0 < valid; e0 += 128) {
        const float4 v4 = ldg4(f + e0);  // in bounds: ld >= 3E and ld % 4 == 0
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + j;
            if (e >= valid) break;
            const int an = (e >= E) + (e >= 2 * E);
            const int fld = e - an * E;
            if (fld < 5 || !((mask >> an) & 1u)) continue;
            const float cf = an == 0 ? c0 : (an == 1 ? c1 : c2);
            const float sc = cf * sigmoid_exact(vv[j]);
            if (sc > a.thr) {
                const size_t list = (size_t)b * C + (fld - 5);
                const int pos = atomicAdd(cand_count + list, 1);
                if (pos < a.cand_cap) {
                    cand_score[list * a.cand_cap + pos] = sc;
                    cand_index[list * a.cand_cap + pos] = box_base + an;
                }
            }
        }
    }
}

// yolo_head as a standalone op (reference code/yolo3/model.py:344-371): the public function of the
// reference call surface; yolo_eval uses the fused decode kernels above instead.
__global__ void __launch_bounds__(256)
yolo_head_kernel(const float* __restrict__ feats, int ld, long long cells, int gh, int gw, int A, int C,
                 const float* __restrict__ anchors, float in_h, float in_w, float* __restrict__ box_xy,
                 float* __restrict__ box_wh, float* __restrict__ conf, float* __restrict__ cls, float* __restrict__ grid) {
    const int E = C + 5;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cells * A * E) return;
    const int e = (int)(idx % E);
    const long long ca = idx / E;  // cell * A + anchor
    const int an = (int)(ca % A);
    const long long cell = ca / A;
    const int gx = (int)(cell % gw), gy = (int)((cell / gw) % gh);
    const float v = __ldg(feats + cell * ld + (size_t)an * E + e);
    if (e == 0) box_xy[ca * 2] = (sigmoid_exact(v) + (float)gx) / (float)gw;
    else if (e == 1) box_xy[ca * 2 + 1] = (sigmoid_exact(v) + (float)gy) / (float)gh;
    else if (e == 2) box_wh[ca * 2] = expf(v) * __ldg(anchors + an * 2) / in_w;
    else if (e == 3) box_wh[ca * 2 + 1] = expf(v) * __ldg(anchors + an * 2 + 1) / in_h;
    else if (e == 4) conf[ca] = sigmoid_exact(v);
    else if (cls != nullptr) cls[ca * C + (e - 5)] = sigmoid_exact(v);
    if (grid != nullptr && an == 0 && e < 2 && cell < (long long)gh * gw) grid[cell * 2 + e] = e == 0 ? (float)gx : (float)gy;
}

// ---- NMS ---------------------------------------------------------------------------
__device__ __forceinline__ float iou_tf(const float4 bi, const float4 bj) {
    // NonMaxSuppressionV3 IOU (boxes are (y0,x0,y1,x1) with possibly swapped corners)
    const float ymin_i = fminf(bi.x, bi.z), xmin_i = fminf(bi.y, bi.w);
    const float ymax_i = fmaxf(bi.x, bi.z), xmax_i = fmaxf(bi.y, bi.w);
    const float ymin_j = fminf(bj.x, bj.z), xmin_j = fminf(bj.y, bj.w);
    const float ymax_j = fmaxf(bj.x, bj.z), xmax_j = fmaxf(bj.y, bj.w);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iy0 = fmaxf(ymin_i, ymin_j), ix0 = fmaxf(xmin_i, xmin_j);
    const float iy1 = fminf(ymax_i, ymax_j), ix1 = fminf(xmax_i, xmax_j);
    const float inter = fmaxf(iy1 - iy0, 0.0f) * fmaxf(ix1 - ix0, 0.0f);
    return inter / (area_i + area_j - inter);
}

__device__ __forceinline__ unsigned long long nms_key(float score, int idx) {
    unsigned u = __float_as_uint(score);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of fp32 -> u32
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}

constexpr int NMS_THREADS = 128;
constexpr int NMS_FAST_MIN = 1024;   // lists longer than this try the shared-memory subset first
constexpr int NMS_CAP = 512;         // subset capacity: 4 KB of keys + 8 KB of boxes (static smem: keeps 16 CTAs per SM)
constexpr int NMS_TARGET = 256;      // the subset holds at least this many of the best candidates

__device__ __forceinline__ unsigned nms_u32(float score) {
    const unsigned u = __float_as_uint(score);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of fp32 -> u32
}

// block-wide maximum of a 64-bit key (all threads get it)
__device__ __forceinline__ unsigned long long nms_block_max(unsigned long long best, unsigned long long* s_key,
                                                            unsigned long long* s_best) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if (lane == 0) s_key[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = s_key[0];
#pragma unroll
        for (int w = 1; w < NMS_THREADS / 32; ++w) m = s_key[w] > m ? s_key[w] : m;
        *s_best = m;
    }
    __syncthreads();
    const unsigned long long r = *s_best;
    __syncthreads();  // s_key / s_best are rewritten by the next call
    return r;
}

// Greedy NMS as 'repeat: pick the best alive candidate (score desc, index asc), kill
// everything with IoU > thr against it'.  This selects exactly the boxes, in exactly the
// order, of the sequential priority-queue algorithm: a candidate is selected iff no
// earlier-selected box suppresses it, and candidates are visited by (score, index).
//
// Long lists (MAP mode runs with score_threshold = 0: every one of the 10 647 boxes is a candidate of every class,
// reference code/main.py:175) first run the same loop on a SUBSET held in shared memory: all candidates whose score
// is >= a threshold found with a two-level 8-bit histogram so that the subset holds the >= 256 best.  Greedy NMS
// visits candidates in key order, so as long as it stops (max_boxes selected) before the subset is exhausted, the
// rest of the list can never be looked at and the result is exactly that of the full list.  If the subset runs out
// first, the remaining candidates are filtered against everything selected so far and the plain loop continues.
__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(const float* __restrict__ boxes, int total_boxes, float* __restrict__ cand_score,
           const int32_t* __restrict__ cand_index, const int32_t* __restrict__ cand_count, int C, int cand_cap,
           int max_boxes, float iou_thr, float* __restrict__ det, int32_t* __restrict__ det_count,
           int32_t* __restrict__ status) {
    __shared__ unsigned long long s_key[NMS_THREADS / 32];
    __shared__ unsigned long long s_best;
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_sel[4];  // bin, count above the bin, count including the bin, subset size
    __shared__ unsigned long long sub_key[NMS_CAP];
    __shared__ float4 sub_box[NMS_CAP];
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t list = (size_t)b * C + c;
    int n = cand_count[list];
    if (n > cand_cap) {
        if (threadIdx.x == 0) atomicExch(status, 1);
        n = cand_cap;
    }
    float* sc = cand_score + list * cand_cap;
    const int32_t* ix = cand_index + list * cand_cap;
    const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * total_boxes;
    float* out = det + list * max_boxes * 6;
    const int tid = threadIdx.x;

    auto emit = [&](int k, unsigned long long best, float4 sel, int sel_idx) {  // thread 0 only
        unsigned u = (unsigned)(best >> 32);
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        out[k * 6 + 0] = sel.x;
        out[k * 6 + 1] = sel.y;
        out[k * 6 + 2] = sel.z;
        out[k * 6 + 3] = sel.w;
        out[k * 6 + 4] = __uint_as_float(u);
        out[k * 6 + 5] = __int_as_float(sel_idx);
    };

    int k = 0;
    if (n > NMS_FAST_MIN) {
        // ---- threshold: largest 16-bit score prefix P with count(prefix >= P) >= NMS_TARGET, if that count fits
        unsigned prefix = 0;
        bool have = false;
        for (int level = 0; level < 2; ++level) {
            for (int i = tid; i < 256; i += NMS_THREADS) s_hist[i] = 0;
            __syncthreads();
            const unsigned hi = s_sel[0];  // bin of level 0 (valid at level 1)
            for (int i = tid; i < n; i += NMS_THREADS) {
                const unsigned u = nms_u32(sc[i]);
                if (level == 0) atomicAdd(&s_hist[u >> 24], 1u);
                else if ((u >> 24) == hi) atomicAdd(&s_hist[(u >> 16) & 0xffu], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned above = level == 0 ? 0u : s_sel[1], cum = above;
                int bin = 255;
                for (; bin > 0; --bin) {
                    if (cum + s_hist[bin] >= (unsigned)NMS_TARGET) break;
                    cum += s_hist[bin];
                }
                s_sel[0] = (unsigned)bin;
                s_sel[1] = cum;                 // strictly above this bin
                s_sel[2] = cum + s_hist[bin];   // including it
            }
            __syncthreads();
            const unsigned bin = s_sel[0], incl = s_sel[2];
            if (level == 0) {
                prefix = bin << 24;
                if (incl <= (unsigned)NMS_CAP) { have = true; break; }
            } else {
                prefix |= bin << 16;
                have = incl <= (unsigned)NMS_CAP;
            }
            __syncthreads();
        }
        if (have) {
            // ---- subset -> shared memory (any order: the arg-max below is order independent)
            if (tid == 0) s_sel[3] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NMS_THREADS) {
                const float sv = sc[i];
                if (nms_u32(sv) >= prefix) {
                    const int id = ix[i];
                    const unsigned pos = atomicAdd(&s_sel[3], 1u);
                    sub_key[pos] = nms_key(sv, id);
                    sub_box[pos] = __ldg(bx + id);
                }
            }
            __syncthreads();
            const int m = (int)s_sel[3];
            float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
            bool have_sel = false;
            bool exhausted = false;
            while (true) {
                unsigned long long best = 0ull;
                for (int i = tid; i < m; i += NMS_THREADS) {
                    const unsigned long long key = sub_key[i];
                    if (key == 0ull) continue;
                    if (have_sel && iou_tf(sub_box[i], sel) > iou_thr) {
                        sub_key[i] = 0ull;
                        continue;
                    }
                    best = key > best ? key : best;
                }
                best = nms_block_max(best, s_key, &s_best);
                if (best == 0ull) { exhausted = true; break; }
                const int sel_idx = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
                sel = __ldg(bx + sel_idx);
                have_sel = true;
                for (int i = tid; i < m; i += NMS_THREADS)
                    if (sub_key[i] == best) sub_key[i] = 0ull;  // the selected one leaves the list
                if (tid == 0) emit(k, best, sel, sel_idx);
                ++k;
                if (k >= max_boxes) break;
                __syncthreads();
            }
            if (!exhausted) {
                if (tid == 0) det_count[list] = k;
                return;
            }
            // ---- subset exhausted before max_boxes: drop it from the list, filter the rest against the selections
            __syncthreads();  // the selections written by thread 0 are visible to the block
            for (int i = tid; i < n; i += NMS_THREADS) {
                const float sv = sc[i];
                bool dead = nms_u32(sv) >= prefix;
                if (!dead) {
                    const float4 bb = __ldg(bx + ix[i]);
                    for (int j = 0; j < k && !dead; ++j) {
                        const float4 sj = make_float4(out[j * 6 + 0], out[j * 6 + 1], out[j * 6 + 2], out[j * 6 + 3]);
                        dead = iou_tf(bb, sj) > iou_thr;
                    }
                }
                if (dead) sc[i] = -INFINITY;
            }
            __syncthreads();
        }
    }

    float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
    int sel_idx = -1;
    while (n > 0 && k < max_boxes) {
        // pass: apply the previous selection's suppression, find the best survivor
        unsigned long long best = 0ull;
        for (int i = tid; i < n; i += NMS_THREADS) {
            float s = sc[i];
            if (s == -INFINITY) continue;
            const int id = ix[i];
            if (sel_idx >= 0) {
                if (id == sel_idx || iou_tf(__ldg(bx + id), sel) > iou_thr) {
                    sc[i] = -INFINITY;
                    continue;
                }
            }
            const unsigned long long key = nms_key(s, id);
            best = key > best ? key : best;
        }
        best = nms_block_max(best, s_key, &s_best);
        if (best == 0ull) break;
        sel_idx = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
        sel = __ldg(bx + sel_idx);
        if (tid == 0) emit(k, best, sel, sel_idx);
        ++k;
    }
    if (tid == 0) det_count[list] = k;
}

__global__ void pack_kernel(const float* __restrict__ det, const int32_t* __restrict__ det_count, int C, int max_boxes,
                            float* __restrict__ out_boxes_f, int32_t* __restrict__ out_boxes_i,
                            float* __restrict__ out_scores, int32_t* __restrict__ out_classes,
                            int32_t* __restrict__ out_count) {
    extern __shared__ int s_off[];  // [C+1]
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < C; ++c) {
            s_off[c] = acc;
            acc += det_count[(size_t)b * C + c];
        }
        s_off[C] = acc;
        out_count[b] = acc;
    }
    __syncthreads();
    const int slots = C * max_boxes;
    const size_t ob = (size_t)b * slots;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        const int c = i / max_boxes, j = i % max_boxes;
        if (j < det_count[(size_t)b * C + c]) {
            const float* d = det + (((size_t)b * C + c) * max_boxes + j) * 6;
            const int o = s_off[c] + j;
            for (int q = 0; q < 4; ++q) {
                out_boxes_f[(ob + o) * 4 + q] = d[q];
                out_boxes_i[(ob + o) * 4 + q] = (int32_t)d[q];  // tf.cast float->int32 truncates (model.py:490)
            }
            out_scores[ob + o] = d[4];
            out_classes[ob + o] = c;
        }
    }
    __syncthreads();
    for (int i = s_off[C] + threadIdx.x; i < slots; i += blockDim.x) {
        for (int q = 0; q < 4; ++q) {
            out_boxes_f[(ob + i) * 4 + q] = 0.f;
            out_boxes_i[(ob + i) * 4 + q] = 0;
        }
        out_scores[ob + i] = 0.f;
        out_classes[ob + i] = -1;
    }
}

}  // namespace yr

using namespace yr;

extern "C" int yr_decode_filter(const float* const feats[3], const float* image_shapes, const yr_decode_params* p,
                                float* boxes, float* cand_score, int32_t* cand_index, int32_t* cand_count,
                                void* stream) {
    YR_CHECK_ARG(p && feats && image_shapes && boxes && cand_score && cand_index && cand_count, "decode: null pointer");
    YR_CHECK_ARG(p->num_scales >= 1 && p->num_scales <= 3, "decode: num_scales=%d", p->num_scales);
    YR_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= 2048, "decode: num_classes=%d", p->num_classes);
    YR_CHECK_ARG(p->B >= 1 && p->B <= 65535 && p->cand_cap >= 1, "decode: bad B / cand_cap");
    DecodeArgs a;
    a.num_scales = p->num_scales;
    a.C = p->num_classes;
    int off = 0;
    for (int s = 0; s < 3; ++s) {
        a.box_off[s] = off;
        if (s < p->num_scales) {
            YR_CHECK_ARG(feats[s] != nullptr, "decode: feats[%d] is null", s);
            YR_CHECK_ARG(p->ld[s] >= 3 * (p->num_classes + 5), "decode: ld[%d]=%d too small", s, p->ld[s]);
            a.feats[s] = feats[s];
            a.gh[s] = p->grid_h[s];
            a.gw[s] = p->grid_w[s];
            a.ld[s] = p->ld[s];
            off += 3 * p->grid_h[s] * p->grid_w[s];
        } else {
            a.feats[s] = nullptr;
            a.gh[s] = a.gw[s] = a.ld[s] = 1;
        }
        for (int k = 0; k < 3; ++k) {
            a.anchors[s][k][0] = p->anchors[s][k][0];
            a.anchors[s][k][1] = p->anchors[s][k][1];
        }
    }
    a.box_off[3] = off;
    a.total_boxes = off;
    a.cand_cap = p->cand_cap;
    a.in_h = (float)p->input_h;
    a.in_w = (float)p->input_w;
    a.thr = p->score_threshold;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)p->B * p->num_classes, s) != cudaSuccess) {
        set_error("decode: memset failed");
        return YR_ERR_CUDA;
    }
    const size_t smem = (size_t)DEC_BOXES * p->num_classes * sizeof(float);
    if (smem > 48 * 1024) {
        static DeviceOnce attr_once;  // function attributes are per device
        bool& attr = attr_once.cur();
        if (!attr) {
            cudaFuncSetAttribute(decode_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr = true;
        }
        YR_CHECK_ARG(smem <= 200 * 1024, "decode: num_classes too large");
    }
    bool aligned = ((uintptr_t)boxes % 16) == 0;
    for (int k = 0; k < p->num_scales; ++k) aligned = aligned && ((uintptr_t)feats[k] % 16 == 0) && (p->ld[k] % 4 == 0);
    if (p->score_threshold >= 0.01f && aligned) {
        const int total_cells = a.total_boxes / 3;
        dim3 grid(cdiv(total_cells, 8), p->B);
        decode_sparse_kernel<<<grid, 256, 0, s>>>(a, image_shapes, boxes, cand_score, cand_index, cand_count, total_cells);
    } else {
        // dense variant (MAP mode runs with score_threshold = 0, reference code/main.py:175): CTA-level
        // compaction keeps the atomics at one per (CTA, class)
        dim3 grid(cdiv(a.total_boxes, DEC_BOXES), p->B);
        decode_filter_kernel<<<grid, DEC_THREADS, smem, s>>>(a, image_shapes, boxes, cand_score, cand_index, cand_count);
    }
    YR_CHECK_LAUNCH("decode_filter");
    return YR_OK;
}

extern "C" int yr_nms_classwise(const float* boxes, int total_boxes, float* cand_score, const int32_t* cand_index,
                                const int32_t* cand_count, int B, int num_classes, int cand_cap, int max_boxes,
                                float iou_threshold, float* det, int32_t* det_count, int32_t* status, void* stream) {
    YR_CHECK_ARG(boxes && cand_score && cand_index && cand_count && det && det_count && status, "nms: null pointer");
    YR_CHECK_ARG(B >= 1 && B <= 65535 && num_classes >= 1 && max_boxes >= 1 && cand_cap >= 1, "nms: bad sizes");
    YR_CHECK_ARG(((uintptr_t)boxes) % 16 == 0, "nms: boxes must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(status, 0, sizeof(int32_t), s) != cudaSuccess) {
        set_error("nms: memset failed");
        return YR_ERR_CUDA;
    }
    dim3 grid(num_classes, B);
    nms_kernel<<<grid, NMS_THREADS, 0, s>>>(boxes, total_boxes, cand_score, cand_index, cand_count, num_classes, cand_cap,
                                            max_boxes, iou_threshold, det, det_count, status);
    YR_CHECK_LAUNCH("nms");
    return YR_OK;
}

extern "C" int yr_pack_detections(const float* det, const int32_t* det_count, int B, int num_classes, int max_boxes,
                                  float* out_boxes_f, int32_t* out_boxes_i, float* out_scores, int32_t* out_classes,
                                  int32_t* out_count, void* stream) {
    YR_CHECK_ARG(det && det_count && out_boxes_f && out_boxes_i && out_scores && out_classes && out_count,
                 "pack: null pointer");
    YR_CHECK_ARG(B >= 1 && num_classes >= 1 && max_boxes >= 1, "pack: bad sizes");
    pack_kernel<<<B, 128, (num_classes + 1) * sizeof(int), (cudaStream_t)stream>>>(
        det, det_count, num_classes, max_boxes, out_boxes_f, out_boxes_i, out_scores, out_classes, out_count);
    YR_CHECK_LAUNCH("pack");
    return YR_OK;
}

extern "C" int yr_yolo_head(const float* feats, int ld, int B, int gh, int gw, int A, int C, const float* anchors_dev,
                            int input_h, int input_w, float* box_xy, float* box_wh, float* box_confidence,
                            float* box_class_probs, float* grid, void* stream) {
    YR_CHECK_ARG(feats && anchors_dev && box_xy && box_wh && box_confidence, "yolo_head: null pointer");
    YR_CHECK_ARG(B >= 1 && gh >= 1 && gw >= 1 && A >= 1 && C >= 1 && ld >= A * (C + 5), "yolo_head: bad sizes");
    const long long cells = (long long)B * gh * gw;
    const long long total = cells * A * (C + 5);
    YR_CHECK_ARG(total < (1ll << 40), "yolo_head: too many elements");
    yolo_head_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        feats, ld, cells, gh, gw, A, C, anchors_dev, (float)input_h, (float)input_w, box_xy, box_wh, box_confidence,
        box_class_probs, grid);
    YR_CHECK_LAUNCH("yolo_head");
    return YR_OK;
}
