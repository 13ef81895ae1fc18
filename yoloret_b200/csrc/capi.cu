// C-ABI glue: error state, plan executor, letterbox pre-process.
#include "yr_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace yr {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("YR_PDL");
        return e != nullptr && atoi(e) != 0;  // opt-in: measured on B200, no gain inside the captured graph
    }();
    return on;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// letterbox_image (reference code/yolo3/utils.py:67-83) fused with the u8 -> [0,1] scaling of
// tf.io.decode_image(dtype=float32) (code/yolo.py:106): bilinear, half-pixel centres, no
// antialias (tf.image.resize defaults), zero padding outside [dy,dy+nh) x [dx,dx+nw).
__global__ void __launch_bounds__(256)
letterbox_kernel(const uint8_t* __restrict__ src, int ih, int iw, float* __restrict__ dst, int h, int w, int nh, int nw,
                 int dy, int dx) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= h * w) return;
    const int y = idx / w, x = idx % w;
    float r = 0.f, g = 0.f, bl = 0.f;
    const int yy = y - dy, xx = x - dx;
    if (yy >= 0 && yy < nh && xx >= 0 && xx < nw) {
        const float sy = (float)ih / (float)nh, sx = (float)iw / (float)nw;
        float fy = ((float)yy + 0.5f) * sy - 0.5f;
        float fx = ((float)xx + 0.5f) * sx - 0.5f;
        fy = fmaxf(fy, 0.0f);
        fx = fmaxf(fx, 0.0f);
        int y0 = (int)fy, x0 = (int)fx;
        y0 = min(y0, ih - 1);
        x0 = min(x0, iw - 1);
        const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float k = 1.0f / 255.0f;
        auto px = [&](int yq, int xq, int c) { return (float)__ldg(src + ((size_t)yq * iw + xq) * 3 + c) * k; };
        auto lerp2 = [&](int c) {
            const float top = px(y0, x0, c) + (px(y0, x1, c) - px(y0, x0, c)) * lx;
            const float bot = px(y1, x0, c) + (px(y1, x1, c) - px(y1, x0, c)) * lx;
            return top + (bot - top) * ly;
        };
        r = lerp2(0);
        g = lerp2(1);
        bl = lerp2(2);
    }
    dst[(size_t)idx * 3 + 0] = r;
    dst[(size_t)idx * 3 + 1] = g;
    dst[(size_t)idx * 3 + 2] = bl;
}

}  // namespace yr

using namespace yr;

extern "C" int yr_version(void) { return 100; }
extern "C" const char* yr_last_error(void) { return g_err; }
extern "C" int yr_sizeof_op(void) { return (int)sizeof(yr_op); }

extern "C" int yr_dw_se_slots(const yr_op* op) {
    if (!op || op->kind != YR_OP_DW || op->C <= 0 || op->C % 4 || (op->k != 3 && op->k != 5) ||
        (op->stride != 1 && op->stride != 2) || op->Ho <= 0 || op->Wo <= 0) {
        set_error("dw_se_slots: not a valid depthwise op");
        return YR_ERR_INVALID;
    }
    return dw_se_slots(*op);
}

extern "C" int yr_run_ops(const yr_op* ops, int n_ops, void* stream) {
    YR_CHECK_ARG(ops != nullptr || n_ops == 0, "run_ops: null ops");
    cudaStream_t s = (cudaStream_t)stream;
    for (int i = 0; i < n_ops; ++i) {
        const yr_op& op = ops[i];
        int rc;
        switch (op.kind) {
            case YR_OP_STEM: rc = launch_stem(op, s); break;
            case YR_OP_PW:
                if (op.K2 != 0 && op.variant != 3 && op.variant != 4) {
                    set_error("run_ops: op %d: stacked outputs (K2 = %d) need the tensor-memory-A kernel (variant 3 or 4)", i, op.K2);
                    return YR_ERR_INVALID;
                }
                if (op.variant == 4) rc = launch_pw_ts2(op, s);
                else if (op.variant == 3) rc = launch_pw_ts(op, s);
                else rc = (op.variant == 2 || (op.variant == 0 && op.w_tc != nullptr)) ? launch_pw_tc(op, s) : launch_pw(op, s);
                break;
            case YR_OP_DW: rc = launch_dw(op, s); break;
            case YR_OP_RESAMPLE: rc = launch_resample(op, s); break;
            case YR_OP_RFCR: rc = launch_rfcr(op, s); break;
            case YR_OP_SE: rc = launch_se(op, s); break;
            case YR_OP_SE_FC: rc = launch_se_fc(op, s); break;
            case YR_OP_DWPW: rc = launch_dwpw(op, s); break;
            default:
                set_error("run_ops: op %d has unknown kind %d", i, op.kind);
                return YR_ERR_INVALID;
        }
        if (rc != YR_OK) {
            char tmp[400];
            strncpy(tmp, g_err, sizeof(tmp) - 1);
            tmp[sizeof(tmp) - 1] = 0;
            set_error("op %d (kind %d): %s", i, op.kind, tmp);
            return rc;
        }
    }
    return YR_OK;
}

extern "C" int yr_letterbox_u8(const uint8_t* src, int ih, int iw, float* dst, int h, int w, int nh, int nw, int dy,
                               int dx, void* stream) {
    YR_CHECK_ARG(src && dst, "letterbox: null pointer");
    YR_CHECK_ARG(ih > 0 && iw > 0 && h > 0 && w > 0 && nh > 0 && nw > 0 && dy >= 0 && dx >= 0 && dy + nh <= h &&
                     dx + nw <= w, "letterbox: bad geometry");
    letterbox_kernel<<<cdiv(h * w, 256), 256, 0, (cudaStream_t)stream>>>(src, ih, iw, dst, h, w, nh, nw, dy, dx);
    YR_CHECK_LAUNCH("letterbox");
    return YR_OK;
}
