// Depthwise kxk convolution (k in {3,5}, stride in {1,2}), NHWC fp32, folded BN + activation.
//
// Replaces DepthwiseConv2D+BatchNormalization+ReLU6/Swish of the Keras MobileNetV2
// blocks (reference code/yolo3/override.py:339), MBConvBlock
// (code/yolo3/efficientnet.py:501-510) and MobilenetSeparableConv2D (code/yolo3/model.py:20-24).
// Per-channel work, 1.7 flop/B: an HBM-bound streaming kernel.  A thread owns 4
// channels (one 128-bit lane of the NHWC pixel) of a TH x TW output patch and walks
// the input rows once, so every input float4 it loads feeds up to k*k/S^2 outputs from
// registers; horizontally/vertically adjacent patches share their halo through L1.
// TF 'SAME' padding is expressed as leading pads (pad_t, pad_l) + bounds checks, which
// reproduces the asymmetric (0,1) padding of stride-2 layers on even inputs.
#include "yr_common.cuh"

namespace yr {

template <int KS, int S, int TH, int TW, int ACT>
__global__ void __launch_bounds__(128)
dw_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ wgt, const float* __restrict__ bias,
          float* __restrict__ out, int ld_out, int H, int W, int C, int Ho, int Wo, int pad_t, int pad_l,
          float* __restrict__ part) {
    constexpr int IN_ROWS = (TH - 1) * S + KS;
    constexpr int IN_COLS = (TW - 1) * S + KS;
    extern __shared__ __align__(16) float s_sum[];  // [C], only when part != nullptr
    pdl_wait();
    pdl_launch_dependents();
    const int C4 = C >> 2;
    const int wtiles = (Wo + TW - 1) / TW;
    const int item = blockIdx.x * 128 + threadIdx.x;
    const bool active = item < wtiles * C4;
    float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);  // sum of this thread's outputs (squeeze-excite mean)
    const int c = active ? (item % C4) * 4 : 0;
    if (active) {
    const int wo0 = (item / C4) * TW;
    const int ho0 = blockIdx.y * TH;
    const int b = blockIdx.z;

    float4 w[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) w[i] = ldg4(wgt + (size_t)i * C + c);

    const float4 bv = ldg4(bias + c);
    float4 acc[TH][TW];
#pragma unroll
    for (int t = 0; t < TH; ++t)
#pragma unroll
        for (int o = 0; o < TW; ++o) acc[t][o] = bv;  // accumulate on top of the folded-BN bias (as dw_tma_kernel)

    const int hi0 = ho0 * S - pad_t;
    const int wi0 = wo0 * S - pad_l;
    const float* inb = in + (size_t)b * H * W * ld_in + c;

#pragma unroll
    for (int r = 0; r < IN_ROWS; ++r) {
        const int hi = hi0 + r;
        if (hi < 0 || hi >= H) continue;
        const float* row = inb + (size_t)hi * W * ld_in;
        float4 x[IN_COLS];
#pragma unroll
        for (int j = 0; j < IN_COLS; ++j) {
            const int wi = wi0 + j;
            x[j] = (wi >= 0 && wi < W) ? ldg4(row + (size_t)wi * ld_in) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            const int kh = r - t * S;
            if (kh < 0 || kh >= KS) continue;
#pragma unroll
            for (int kw = 0; kw < KS; ++kw) {
                const float4 ww = w[kh * KS + kw];
#pragma unroll
                for (int o = 0; o < TW; ++o) {
                    fma4(acc[t][o], x[o * S + kw], ww);
                }
            }
        }
    }

#pragma unroll
    for (int t = 0; t < TH; ++t) {
        const int ho = ho0 + t;
        if (ho >= Ho) continue;
#pragma unroll
        for (int o = 0; o < TW; ++o) {
            const int wo = wo0 + o;
            if (wo >= Wo) continue;
            float4 v;
            v.x = apply_act<ACT>(acc[t][o].x);
            v.y = apply_act<ACT>(acc[t][o].y);
            v.z = apply_act<ACT>(acc[t][o].z);
            v.w = apply_act<ACT>(acc[t][o].w);
            st4(out + (((size_t)b * Ho + ho) * Wo + wo) * ld_out + c, v);
            ps.x += v.x; ps.y += v.y; ps.z += v.z; ps.w += v.w;
        }
    }
    }  // active
    if (part != nullptr) {
        // Squeeze-excite 'Mean' (reference code/yolo3/efficientnet.py:391-403,419) fused here: a
        // deterministic per-CTA channel sum (threads that share a channel group add in rank order),
        // one [C] slot per CTA; se_fc_kernel adds the slots in index order.  No atomics: the result
        // is bit-identical run to run, which the bit-exact NMS tests downstream rely on.
        for (int i = threadIdx.x; i < C; i += 128) s_sum[i] = 0.0f;
        __syncthreads();
        const int G = (128 + C4 - 1) / C4;
        for (int g = 0; g < G; ++g) {
            if (active && (int)threadIdx.x / C4 == g) {
                float4 a = *reinterpret_cast<float4*>(s_sum + c);
                a.x += ps.x; a.y += ps.y; a.z += ps.z; a.w += ps.w;
                *reinterpret_cast<float4*>(s_sum + c) = a;
            }
            __syncthreads();
        }
        float* dst = part + (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * C;
        for (int i = threadIdx.x; i < C4; i += 128) st4(dst + i * 4, *reinterpret_cast<float4*>(s_sum + i * 4));
    }
}

template <int KS, int S, int TH, int TW, int ACT>
static int launch_dw_cfg(const yr_op& op, cudaStream_t s) {
    const int items = cdiv(op.Wo, TW) * (op.C / 4);
    dim3 grid(cdiv(items, 128), cdiv(op.Ho, TH), op.B);
    const size_t smem = op.aux ? (size_t)op.C * sizeof(float) : 0;
    if (launch_pdl(dw_kernel<KS, S, TH, TW, ACT>, grid, dim3(128), smem, s, (const float*)op.in, op.ld_in, op.w, op.bias,
                   (float*)op.out, op.ld_out, op.H, op.W, op.C, op.Ho, op.Wo, op.pad_t, op.pad_l, op.aux) != cudaSuccess) {
        set_error("dw: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    return YR_OK;
}

static void dw_tile(int k, int stride, int& th, int& tw) {
    th = (k == 3) ? 2 : 1;
    tw = (k == 5 && stride == 2) ? 2 : 4;
}

// Number of [C] partial-sum slots per image the fused squeeze-excite sum of this DW op writes.
int dw_se_slots(const yr_op& op_in) {
    yr_op op = op_in;
    if (!op.aux) op.aux = reinterpret_cast<float*>(uintptr_t(16));  // "will have a partial-sum buffer": same kernel choice as the launch
    if (dw_uses_tma(op)) return dw_tma_se_slots(op);
    int th, tw;
    dw_tile(op.k, op.stride, th, tw);
    return cdiv(cdiv(op.Wo, tw) * (op.C / 4), 128) * cdiv(op.Ho, th);
}

template <int ACT>
static int launch_dw_act(const yr_op& op, cudaStream_t s) {
    if (op.k == 3 && op.stride == 1) return launch_dw_cfg<3, 1, 2, 4, ACT>(op, s);
    if (op.k == 3 && op.stride == 2) return launch_dw_cfg<3, 2, 2, 4, ACT>(op, s);
    if (op.k == 5 && op.stride == 1) return launch_dw_cfg<5, 1, 1, 4, ACT>(op, s);
    if (op.k == 5 && op.stride == 2) return launch_dw_cfg<5, 2, 1, 2, ACT>(op, s);
    set_error("dw: unsupported k=%d stride=%d", op.k, op.stride);
    return YR_ERR_UNSUPPORTED;
}

int launch_dw(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "dw: null pointer");
    YR_CHECK_ARG(op.C > 0 && op.C % 4 == 0, "dw: C=%d must be a multiple of 4", op.C);
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= op.C && op.ld_out % 4 == 0, "dw: bad ld");
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w | (uintptr_t)op.bias) % 16 == 0,
                 "dw: pointers must be 16-byte aligned");
    YR_CHECK_ARG(op.B <= 65535 && cdiv(op.Ho, 1) <= 65535, "dw: grid too large");
    YR_CHECK_ARG(!op.aux || ((uintptr_t)op.aux % 16 == 0 && op.C <= 12288), "dw: bad squeeze-excite partial buffer");
    if (dw_uses_tma(op)) return launch_dw_tma(op, s);
    switch (op.act) {
        case YR_ACT_NONE: return launch_dw_act<YR_ACT_NONE>(op, s);
        case YR_ACT_RELU6: return launch_dw_act<YR_ACT_RELU6>(op, s);
        case YR_ACT_SWISH: return launch_dw_act<YR_ACT_SWISH>(op, s);
    }
    set_error("dw: unknown activation %d", op.act);
    return YR_ERR_INVALID;
}

}  // namespace yr
