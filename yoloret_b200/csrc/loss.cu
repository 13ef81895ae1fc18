// Fused YoloLoss forward + backward for one scale.
//
// Reference: YoloLoss.call (code/yolo3/model.py:607-671) + do_giou_calculate
// (code/yolo3/utils.py:9-53) + yolo_head(calc_loss=True) (model.py:344-369).  TensorFlow
// materialises iou[B,H,W,A,N_true] (model.py:644-648) and back-propagates through ~60
// element-wise ops; here one warp owns one (cell, anchor): the batch-wide true boxes
// stream through shared memory for the ignore mask (never stored), and the analytic
// gradient w.r.t. the raw logits is written in the same pass.  Gradient conventions are
// TensorFlow's: maximum/minimum send the gradient to the first argument on ties
// (x >= y / x <= y), maximum(0, v) passes it iff v > 0, divide_no_nan has zero gradient
// where the denominator is 0, and the ignore mask (a cast of a comparison) is a constant.
// Reductions are two-stage with a fixed order, so the loss is run-to-run deterministic.
#include "yr_common.cuh"
#include <math.h>

namespace yr {

constexpr int LOSS_WARPS = 8;
constexpr int LOSS_ITER = 4;       // (cell, anchor) items per warp
constexpr int LOSS_CHUNK = 1024;   // true boxes staged per pass

__global__ void __launch_bounds__(256)
gather_true_kernel(const float* __restrict__ y_true, long long items, int A, int C, int ld_true,
                   float* __restrict__ true_boxes, int32_t* __restrict__ n_true, int max_true) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= items) return;
    const long long cell = it / A;
    const int a = (int)(it % A);
    const float* y = y_true + cell * ld_true + (size_t)a * (5 + C);
    if (y[4] == 0.0f) return;  // tf.cast(object_mask, 'bool'), model.py:641
    const float x = y[0], yy = y[1], w = y[2], h = y[3];
    float4 bx;  // (ymin, xmin, ymax, xmax) clipped to [0,1], model.py:635-640
    bx.x = fminf(fmaxf(yy - h / 2.f, 0.f), 1.f);
    bx.y = fminf(fmaxf(x - w / 2.f, 0.f), 1.f);
    bx.z = fminf(fmaxf(yy + h / 2.f, 0.f), 1.f);
    bx.w = fminf(fmaxf(x + w / 2.f, 0.f), 1.f);
    const int pos = atomicAdd(n_true, 1);
    if (pos < max_true) *reinterpret_cast<float4*>(true_boxes + 4 * (size_t)pos) = bx;
}

__device__ __forceinline__ float div_no_nan(float a, float b) { return b == 0.0f ? 0.0f : a / b; }

__device__ __forceinline__ float iou_only(const float4 p, const float4 q) {
    const float pa = fmaxf(0.f, p.w - p.y) * fmaxf(0.f, p.z - p.x);
    const float qa = fmaxf(0.f, q.w - q.y) * fmaxf(0.f, q.z - q.x);
    const float iw = fmaxf(0.f, fminf(p.w, q.w) - fmaxf(p.y, q.y));
    const float ih = fmaxf(0.f, fminf(p.z, q.z) - fmaxf(p.x, q.x));
    const float I = iw * ih;
    return div_no_nan(I, pa + qa - I);
}

// GIoU(p, q) and d GIoU / d p (p = pred (ymin,xmin,ymax,xmax)).
__device__ __forceinline__ float giou_fwd_bwd(const float4 p, const float4 q, float4& dp) {
    const float pw_raw = p.w - p.y, ph_raw = p.z - p.x;
    const float pw = fmaxf(0.f, pw_raw), ph = fmaxf(0.f, ph_raw);
    const float qw = fmaxf(0.f, q.w - q.y), qh = fmaxf(0.f, q.z - q.x);
    const float pa = pw * ph, qa = qw * qh;
    const float iy0 = fmaxf(p.x, q.x), ix0 = fmaxf(p.y, q.y), iy1 = fminf(p.z, q.z), ix1 = fminf(p.w, q.w);
    const float iw_raw = ix1 - ix0, ih_raw = iy1 - iy0;
    const float iw = fmaxf(0.f, iw_raw), ih = fmaxf(0.f, ih_raw);
    const float I = iw * ih;
    const float U = pa + qa - I;
    const float iou = div_no_nan(I, U);
    const float ey0 = fminf(p.x, q.x), ex0 = fminf(p.y, q.y), ey1 = fmaxf(p.z, q.z), ex1 = fmaxf(p.w, q.w);
    const float ew_raw = ex1 - ex0, eh_raw = ey1 - ey0;
    const float ew = fmaxf(0.f, ew_raw), eh = fmaxf(0.f, eh_raw);
    const float E = ew * eh;
    const float D = E - U;
    const float giou = iou - div_no_nan(D, E);
    // backward, upstream 1
    float dI = 0.f, dU = 0.f, dE = 0.f;
    if (U != 0.f) { dI += 1.f / U; dU += -I / (U * U); }
    if (E != 0.f) {
        const float dr = -1.f;
        const float dD = dr / E;
        dE += dD - dr * D / (E * E);
        dU += -dD;
    }
    const float dpa = dU;
    dI += -dU;
    const float diw = (iw_raw > 0.f) ? dI * ih : 0.f;
    const float dih = (ih_raw > 0.f) ? dI * iw : 0.f;
    const float dew = (ew_raw > 0.f) ? dE * eh : 0.f;
    const float deh = (eh_raw > 0.f) ? dE * ew : 0.f;
    const float dpw = (pw_raw > 0.f) ? dpa * ph : 0.f;
    const float dph = (ph_raw > 0.f) ? dpa * pw : 0.f;
    float dy0 = 0.f, dx0 = 0.f, dy1 = 0.f, dx1 = 0.f;
    // intersection corners: max(p,q) -> p if p >= q ; min(p,q) -> p if p <= q
    if (p.y >= q.y) dx0 += -diw;
    if (p.w <= q.w) dx1 += diw;
    if (p.x >= q.x) dy0 += -dih;
    if (p.z <= q.z) dy1 += dih;
    // enclosing corners: min(p,q) -> p if p <= q ; max(p,q) -> p if p >= q
    if (p.y <= q.y) dx0 += -dew;
    if (p.w >= q.w) dx1 += dew;
    if (p.x <= q.x) dy0 += -deh;
    if (p.z >= q.z) dy1 += deh;
    // own area
    dx0 += -dpw; dx1 += dpw; dy0 += -dph; dy1 += dph;
    dp = make_float4(dy0, dx0, dy1, dx1);
    return giou;
}

__device__ __forceinline__ float bce_logits(float x, float z) {  // tf.nn.sigmoid_cross_entropy_with_logits
    return fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
}

struct LossArgs {
    const float* logits;
    const float* y_true;
    const float* true_boxes;
    const int32_t* n_true;
    float* dlogits;
    float* partials;  // [gridDim.x][4]
    long long items;  // B*gh*gw*A
    int gh, gw, A, C, ld_logits, ld_true, max_true;
    float anchors[3][2];
    float in_h, in_w, ignore_thresh, inv_b;
};

__global__ void __launch_bounds__(LOSS_WARPS * 32)
loss_kernel(LossArgs a) {
    __shared__ float4 s_true[LOSS_CHUNK];
    __shared__ float s_part[LOSS_WARPS][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.C, A = a.A, E = 5 + C;
    int nt = *a.n_true;
    nt = nt < a.max_true ? nt : a.max_true;

    long long item[LOSS_ITER];
    float4 pbox[LOSS_ITER];
    float best[LOSS_ITER];
    float pxy_wh[LOSS_ITER][2];  // pred w, h (for d/dt_wh)
    float sxy[LOSS_ITER][2];     // sigmoid(tx), sigmoid(ty)
    const long long base = ((long long)blockIdx.x * LOSS_WARPS + warp) * LOSS_ITER;
#pragma unroll
    for (int it = 0; it < LOSS_ITER; ++it) {
        item[it] = base + it;
        best[it] = -INFINITY;
        pbox[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        pxy_wh[it][0] = pxy_wh[it][1] = sxy[it][0] = sxy[it][1] = 0.f;
        if (item[it] < a.items) {
            const long long cell = item[it] / A;
            const int an = (int)(item[it] % A);
            const int gx = (int)(cell % a.gw), gy = (int)((cell / a.gw) % a.gh);
            const float* t = a.logits + cell * a.ld_logits + (size_t)an * E;
            const float v = lane < 4 ? __ldg(t + lane) : 0.f;
            const float tx = __shfl_sync(0xffffffffu, v, 0), ty = __shfl_sync(0xffffffffu, v, 1);
            const float tw = __shfl_sync(0xffffffffu, v, 2), th = __shfl_sync(0xffffffffu, v, 3);
            const float sx = 1.f / (1.f + expf(-tx)), sy = 1.f / (1.f + expf(-ty));
            const float px = (sx + (float)gx) / (float)a.gw, py = (sy + (float)gy) / (float)a.gh;
            const float pw = expf(tw) * a.anchors[an][0] / a.in_w, ph = expf(th) * a.anchors[an][1] / a.in_h;
            pbox[it] = make_float4(py - ph / 2.f, px - pw / 2.f, py + ph / 2.f, px + pw / 2.f);
            pxy_wh[it][0] = pw; pxy_wh[it][1] = ph;
            sxy[it][0] = sx; sxy[it][1] = sy;
        }
    }
    // ignore mask: best IoU against every true box of the batch (model.py:643-649)
    for (int c0 = 0; c0 < nt; c0 += LOSS_CHUNK) {
        const int cn = min(LOSS_CHUNK, nt - c0);
        __syncthreads();
        for (int j = threadIdx.x; j < cn; j += blockDim.x)
            s_true[j] = __ldg(reinterpret_cast<const float4*>(a.true_boxes) + c0 + j);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < LOSS_ITER; ++it)
            for (int j = lane; j < cn; j += 32) best[it] = fmaxf(best[it], iou_only(pbox[it], s_true[j]));
    }
    float l_giou = 0.f, l_conf = 0.f, l_cls = 0.f, l_ign = 0.f;
#pragma unroll
    for (int it = 0; it < LOSS_ITER; ++it) {
        if (item[it] >= a.items) continue;
        float bi = best[it];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bi = fmaxf(bi, __shfl_xor_sync(0xffffffffu, bi, o));
        const float ignore = bi < a.ignore_thresh ? 1.f : 0.f;
        const long long cell = item[it] / A;
        const int an = (int)(item[it] % A);
        const float* t = a.logits + cell * a.ld_logits + (size_t)an * E;
        const float* y = a.y_true + cell * a.ld_true + (size_t)an * E;
        float* d = a.dlogits ? a.dlogits + cell * a.ld_logits + (size_t)an * E : nullptr;
        const float yv = lane < 5 ? __ldg(y + lane) : 0.f;
        const float obj = __shfl_sync(0xffffffffu, yv, 4);
        float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
        float giou = 0.f;
        if (obj != 0.f) {
            const float qx = __shfl_sync(0xffffffffu, yv, 0), qy = __shfl_sync(0xffffffffu, yv, 1);
            const float qw = __shfl_sync(0xffffffffu, yv, 2), qh = __shfl_sync(0xffffffffu, yv, 3);
            float4 q;
            q.x = fminf(fmaxf(qy - qh / 2.f, 0.f), 1.f);
            q.y = fminf(fmaxf(qx - qw / 2.f, 0.f), 1.f);
            q.z = fminf(fmaxf(qy + qh / 2.f, 0.f), 1.f);
            q.w = fminf(fmaxf(qx + qw / 2.f, 0.f), 1.f);
            giou = giou_fwd_bwd(pbox[it], q, dp);
        }
        float cls = 0.f;
        for (int e = lane; e < E; e += 32) {
            const float x = __ldg(t + e);
            float g = 0.f;
            if (e >= 5) {
                const float z = __ldg(y + e);
                if (obj != 0.f) {
                    cls += obj * bce_logits(x, z);
                    g = obj * (1.f / (1.f + expf(-x)) - z) * a.inv_b;
                }
            } else if (e == 4) {
                const float ce = bce_logits(x, obj);
                const float wgt = obj + (1.f - obj) * ignore;
                l_conf += obj * ce + (1.f - obj) * ce * ignore;
                g = wgt * (1.f / (1.f + expf(-x)) - obj) * a.inv_b;
            } else if (obj != 0.f) {
                // d(obj*(1-giou))/dt = -obj * dGIoU/dbox * dbox/dt
                const float up = -obj * a.inv_b;
                if (e == 0) g = up * (dp.y + dp.w) * sxy[it][0] * (1.f - sxy[it][0]) / (float)a.gw;
                if (e == 1) g = up * (dp.x + dp.z) * sxy[it][1] * (1.f - sxy[it][1]) / (float)a.gh;
                if (e == 2) g = up * (dp.w - dp.y) * 0.5f * pxy_wh[it][0];
                if (e == 3) g = up * (dp.z - dp.x) * 0.5f * pxy_wh[it][1];
            }
            if (d) d[e] = g;
        }
        if (d && an == A - 1)  // zero the pad columns of a padded cell row
            for (int e = A * E + lane; e < a.ld_logits; e += 32) a.dlogits[cell * a.ld_logits + e] = 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cls += __shfl_xor_sync(0xffffffffu, cls, o);
        l_cls += cls;  // every lane holds the warp sum; lane 0's copy is used
        if (lane == 0) {
            l_giou += obj * (1.f - giou);
            l_ign += ignore;
        }
    }
    // l_conf lives on lane 4 % 32 == 4 only
    l_conf = __shfl_sync(0xffffffffu, l_conf, 4);
    if (lane == 0) {
        s_part[warp][0] = l_giou;
        s_part[warp][1] = l_conf;
        s_part[warp][2] = l_cls;
        s_part[warp][3] = l_ign;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (int w = 0; w < LOSS_WARPS; ++w) s += s_part[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
}

__global__ void loss_reduce_kernel(const float* __restrict__ partials, int nblocks, float inv_b,
                                   float* __restrict__ loss_parts) {
    __shared__ double s[4][256];
    const int q = threadIdx.x & 3, t = threadIdx.x >> 2;  // 1024 threads: 256 per component
    double acc = 0.0;
    for (int i = t; i < nblocks; i += 256) acc += (double)partials[(size_t)i * 4 + q];
    s[q][t] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) s[q][t] += s[q][t + o];
        __syncthreads();
    }
    if (t == 0) loss_parts[q] = (float)(q < 3 ? s[q][0] * (double)inv_b : s[q][0]);
}

static long long loss_blocks(const yr_loss_params* p) {
    const long long items = (long long)p->B * p->gh * p->gw * p->A;
    const long long per_block = LOSS_WARPS * LOSS_ITER;
    return (items + per_block - 1) / per_block;
}

}  // namespace yr

using namespace yr;

extern "C" int64_t yr_yolo_loss_workspace(const yr_loss_params* p) {
    if (!p) return 0;
    return (int64_t)loss_blocks(p) * 4 * (int64_t)sizeof(float);
}

extern "C" int yr_yolo_loss_gather_true(const float* y_true, const yr_loss_params* p, float* true_boxes,
                                        int32_t* n_true, void* stream) {
    YR_CHECK_ARG(y_true && p && true_boxes && n_true, "loss_gather: null pointer");
    YR_CHECK_ARG(p->A >= 1 && p->A <= 3 && p->C >= 1 && p->ld_true >= p->A * (5 + p->C), "loss_gather: bad A/C/ld");
    YR_CHECK_ARG(((uintptr_t)true_boxes) % 16 == 0, "loss_gather: true_boxes must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(n_true, 0, sizeof(int32_t), s) != cudaSuccess) {
        set_error("loss_gather: memset failed");
        return YR_ERR_CUDA;
    }
    const long long items = (long long)p->B * p->gh * p->gw * p->A;
    gather_true_kernel<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(y_true, items, p->A, p->C, p->ld_true, true_boxes,
                                                                        n_true, p->max_true);
    YR_CHECK_LAUNCH("loss_gather");
    return YR_OK;
}

extern "C" int yr_yolo_loss(const float* logits, const float* y_true, const float* true_boxes, const int32_t* n_true,
                            const yr_loss_params* p, float* loss_parts, float* dlogits, void* workspace,
                            int64_t workspace_bytes, void* stream) {
    YR_CHECK_ARG(logits && y_true && true_boxes && n_true && p && loss_parts && workspace, "loss: null pointer");
    YR_CHECK_ARG(p->A >= 1 && p->A <= 3 && p->C >= 1, "loss: bad A/C");
    YR_CHECK_ARG(p->ld_logits >= p->A * (5 + p->C) && p->ld_true >= p->A * (5 + p->C), "loss: ld too small");
    YR_CHECK_ARG(((uintptr_t)true_boxes) % 16 == 0, "loss: true_boxes must be 16-byte aligned");
    if (workspace_bytes < yr_yolo_loss_workspace(p)) {
        set_error("loss: workspace %lld < required %lld", (long long)workspace_bytes, (long long)yr_yolo_loss_workspace(p));
        return YR_ERR_WORKSPACE;
    }
    LossArgs a;
    a.logits = logits;
    a.y_true = y_true;
    a.true_boxes = true_boxes;
    a.n_true = n_true;
    a.dlogits = dlogits;
    a.partials = (float*)workspace;
    a.items = (long long)p->B * p->gh * p->gw * p->A;
    a.gh = p->gh; a.gw = p->gw; a.A = p->A; a.C = p->C;
    a.ld_logits = p->ld_logits; a.ld_true = p->ld_true; a.max_true = p->max_true;
    for (int i = 0; i < 3; ++i) { a.anchors[i][0] = p->anchors[i][0]; a.anchors[i][1] = p->anchors[i][1]; }
    a.in_h = (float)p->input_h; a.in_w = (float)p->input_w;
    a.ignore_thresh = p->ignore_thresh;
    a.inv_b = 1.0f / (float)p->B;
    const long long nb = loss_blocks(p);
    YR_CHECK_ARG(nb < (1ll << 31), "loss: too many items");
    cudaStream_t s = (cudaStream_t)stream;
    loss_kernel<<<(unsigned)nb, LOSS_WARPS * 32, 0, s>>>(a);
    YR_CHECK_LAUNCH("loss");
    loss_reduce_kernel<<<1, 1024, 0, s>>>((const float*)workspace, (int)nb, a.inv_b, loss_parts);
    YR_CHECK_LAUNCH("loss_reduce");
    return YR_OK;
}
