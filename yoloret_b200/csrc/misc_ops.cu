// Small streaming kernels of the YOLO-ReT graph: stem conv, nearest/max resampling,
// the fused RFCR fusion and the squeeze-excite gate.  All are HBM/launch bound.
#include "yr_common.cuh"

namespace yr {

// ---------------------------------------------------------------------------------
// Stem: dense 3x3 stride-2 conv with Cin = 3 (K = 27) + folded BN + activation.
// Keras MobileNetV2 Conv1/bn_Conv1/Conv1_relu (reference code/yolo3/override.py:339)
// and the EfficientNet stem (code/yolo3/efficientnet.py:636-645).  u8 input is scaled
// by fp32(1/255) exactly like tf.io.decode_image(dtype=float32) (code/yolo.py:106).
// ---------------------------------------------------------------------------------
template <int ACT, bool U8>
__global__ void __launch_bounds__(256)
stem_kernel(const void* __restrict__ in_, const float* __restrict__ wgt, const float* __restrict__ bias,
            float* __restrict__ out, int ld_out, int B, int H, int W, int Ho, int Wo, int N, int stride, int pad_t,
            int pad_l) {
    extern __shared__ __align__(16) float sw[];  // [27][N]
    for (int i = threadIdx.x; i < 27 * N; i += blockDim.x) sw[i] = wgt[i];
    __syncthreads();
    const int N4 = N >> 2;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * N4;
    if (item >= total) return;
    const int n = (int)(item % N4) * 4;
    const long long p = item / N4;
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int b = (int)(p / ((long long)Wo * Ho));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * stride - pad_t + kh;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int wi = wo * stride - pad_l + kw;
            if (wi < 0 || wi >= W) continue;
            const size_t off = (((size_t)b * H + hi) * W + wi) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                float x;
                if (U8) x = (float)__ldg(reinterpret_cast<const uint8_t*>(in_) + off + ci) * (1.0f / 255.0f);
                else x = __ldg(reinterpret_cast<const float*>(in_) + off + ci);
                const float4 wv = *reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * 3 + ci) * N + n);
                acc.x = fmaf(x, wv.x, acc.x);
                acc.y = fmaf(x, wv.y, acc.y);
                acc.z = fmaf(x, wv.z, acc.z);
                acc.w = fmaf(x, wv.w, acc.w);
            }
        }
    }
    const float4 bv = ldg4(bias + n);
    float4 v;
    v.x = apply_act<ACT>(acc.x + bv.x);
    v.y = apply_act<ACT>(acc.y + bv.y);
    v.z = apply_act<ACT>(acc.z + bv.z);
    v.w = apply_act<ACT>(acc.w + bv.w);
    st4(out + (size_t)p * ld_out + n, v);
}

// v2: one thread owns CPT consecutive output channels of one output pixel, so the 27 input taps are
// loaded once and reused CPT times from registers; weights are shared-memory broadcasts (all lanes of
// a warp with the same channel slice read the same address).  FFMA-bound at ~the HBM time of the layer.
template <int ACT, bool U8, int CPT>
__global__ void __launch_bounds__(256)
stem_kernel_v2(const void* __restrict__ in_, const float* __restrict__ wgt, const float* __restrict__ bias,
               float* __restrict__ out, int ld_out, int B, int H, int W, int Ho, int Wo, int N, int stride, int pad_t,
               int pad_l) {
    extern __shared__ __align__(16) float sw[];  // [27][N]
    for (int i = threadIdx.x; i < 27 * N; i += blockDim.x) sw[i] = wgt[i];
    __syncthreads();
    const int groups = N / CPT;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * groups;
    if (item >= total) return;
    const int n = (int)(item % groups) * CPT;
    const long long p = item / groups;
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int b = (int)(p / ((long long)Wo * Ho));
    float x[27];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho * stride - pad_t + kh;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int wi = wo * stride - pad_l + kw;
            const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
            const size_t off = (((size_t)b * H + (ok ? hi : 0)) * W + (ok ? wi : 0)) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                float v;
                if (U8) v = (float)__ldg(reinterpret_cast<const uint8_t*>(in_) + off + ci) * (1.0f / 255.0f);
                else v = __ldg(reinterpret_cast<const float*>(in_) + off + ci);
                x[(kh * 3 + kw) * 3 + ci] = ok ? v : 0.0f;
            }
        }
    }
    float acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.0f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
#pragma unroll
        for (int j4 = 0; j4 < CPT / 4; ++j4) {
            const float4 wv = *reinterpret_cast<const float4*>(sw + t * N + n + j4 * 4);
            acc[j4 * 4 + 0] = fmaf(x[t], wv.x, acc[j4 * 4 + 0]);
            acc[j4 * 4 + 1] = fmaf(x[t], wv.y, acc[j4 * 4 + 1]);
            acc[j4 * 4 + 2] = fmaf(x[t], wv.z, acc[j4 * 4 + 2]);
            acc[j4 * 4 + 3] = fmaf(x[t], wv.w, acc[j4 * 4 + 3]);
        }
    }
    float* o = out + (size_t)p * ld_out + n;
#pragma unroll
    for (int j4 = 0; j4 < CPT / 4; ++j4) {
        const float4 bv = ldg4(bias + n + j4 * 4);
        float4 v;
        v.x = apply_act<ACT>(acc[j4 * 4 + 0] + bv.x);
        v.y = apply_act<ACT>(acc[j4 * 4 + 1] + bv.y);
        v.z = apply_act<ACT>(acc[j4 * 4 + 2] + bv.z);
        v.w = apply_act<ACT>(acc[j4 * 4 + 3] + bv.w);
        st4(o + j4 * 4, v);
    }
}

template <int ACT, bool U8, int CPT>
static void launch_stem_v2(const yr_op& op, cudaStream_t s) {
    const long long total = (long long)op.B * op.Ho * op.Wo * (op.N / CPT);
    const unsigned grid = (unsigned)((total + 255) / 256);
    const size_t smem = (size_t)27 * op.N * sizeof(float);
    stem_kernel_v2<ACT, U8, CPT><<<grid, 256, smem, s>>>(op.in, op.w, op.bias, (float*)op.out, op.ld_out, op.B, op.H, op.W,
                                                          op.Ho, op.Wo, op.N, op.stride, op.pad_t, op.pad_l);
}

template <int ACT>
static int launch_stem_act(const yr_op& op, cudaStream_t s) {
    const long long total = (long long)op.B * op.Ho * op.Wo * (op.N / 4);
    const unsigned grid = (unsigned)((total + 255) / 256);
    const size_t smem = (size_t)27 * op.N * sizeof(float);
    if (op.N % 12 == 0) {
        if (op.in_is_u8) launch_stem_v2<ACT, true, 12>(op, s);
        else launch_stem_v2<ACT, false, 12>(op, s);
    } else if (op.N % 8 == 0) {
        if (op.in_is_u8) launch_stem_v2<ACT, true, 8>(op, s);
        else launch_stem_v2<ACT, false, 8>(op, s);
    } else if (op.in_is_u8)
        stem_kernel<ACT, true><<<grid, 256, smem, s>>>(op.in, op.w, op.bias, (float*)op.out, op.ld_out, op.B, op.H, op.W,
                                                       op.Ho, op.Wo, op.N, op.stride, op.pad_t, op.pad_l);
    else
        stem_kernel<ACT, false><<<grid, 256, smem, s>>>(op.in, op.w, op.bias, (float*)op.out, op.ld_out, op.B, op.H, op.W,
                                                        op.Ho, op.Wo, op.N, op.stride, op.pad_t, op.pad_l);
    YR_CHECK_LAUNCH("stem");
    return YR_OK;
}

int launch_stem(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "stem: null pointer");
    YR_CHECK_ARG(op.C == 3 && op.k == 3, "stem: needs Cin=3, k=3 (got C=%d k=%d)", op.C, op.k);
    YR_CHECK_ARG(op.N % 4 == 0 && op.N <= 256 && op.ld_out % 4 == 0 && op.ld_out >= op.N, "stem: bad N/ld_out");
    switch (op.act) {
        case YR_ACT_NONE: return launch_stem_act<YR_ACT_NONE>(op, s);
        case YR_ACT_RELU6: return launch_stem_act<YR_ACT_RELU6>(op, s);
        case YR_ACT_SWISH: return launch_stem_act<YR_ACT_SWISH>(op, s);
    }
    set_error("stem: unknown activation %d", op.act);
    return YR_ERR_INVALID;
}

// ---------------------------------------------------------------------------------
// Resample into a channel slice: UpSampling2D() (nearest x2) and MaxPooling2D((s,s))
// (downsample_layer), reference code/yolo3/model.py:139-144,164-166,254,274,307,320.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

template <int MODE>
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int ld_out, int B, int H, int W,
                int C, int Ho, int Wo) {
    const int C4 = C >> 2;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (long long)B * Ho * Wo * C4) return;
    const int c = (int)(item % C4) * 4;
    const long long p = item / C4;
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int b = (int)(p / ((long long)Wo * Ho));
    const float* inb = in + (size_t)b * H * W * ld_in + c;
    float4 v;
    if (MODE == YR_UP2) {
        v = ldg4(inb + ((size_t)(ho >> 1) * W + (wo >> 1)) * ld_in);
    } else {
        constexpr int P = MODE == YR_POOL2 ? 2 : 4;
        v = ldg4(inb + ((size_t)(ho * P) * W + wo * P) * ld_in);
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j < P; ++j)
                if (i | j) v = max4(v, ldg4(inb + ((size_t)(ho * P + i) * W + wo * P + j) * ld_in));
    }
    st4(out + (size_t)p * ld_out + c, v);
}

int launch_resample(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out, "resample: null pointer");
    YR_CHECK_ARG(op.C % 4 == 0 && op.ld_in % 4 == 0 && op.ld_out % 4 == 0, "resample: C/ld must be multiples of 4");
    const long long total = (long long)op.B * op.Ho * op.Wo * (op.C / 4);
    const unsigned grid = (unsigned)((total + 255) / 256);
    const float* in = (const float*)op.in;
    float* out = (float*)op.out;
    if (op.mode == YR_UP2) {
        YR_CHECK_ARG(op.Ho == op.H * 2 && op.Wo == op.W * 2, "resample up2: bad output size");
        resample_kernel<YR_UP2><<<grid, 256, 0, s>>>(in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
    } else if (op.mode == YR_POOL2) {
        YR_CHECK_ARG(op.Ho == op.H / 2 && op.Wo == op.W / 2, "resample pool2: bad output size");
        resample_kernel<YR_POOL2><<<grid, 256, 0, s>>>(in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
    } else if (op.mode == YR_POOL4) {
        YR_CHECK_ARG(op.Ho == op.H / 4 && op.Wo == op.W / 4, "resample pool4: bad output size");
        resample_kernel<YR_POOL4><<<grid, 256, 0, s>>>(in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
    } else {
        set_error("resample: unknown mode %d", op.mode);
        return YR_ERR_INVALID;
    }
    YR_CHECK_LAUNCH("resample");
    return YR_OK;
}

// ---------------------------------------------------------------------------------
// RFCR fusion (reference rfcr_module + WeightedSum, code/yolo3/model.py:117-157, and the
// stride-4 MaxPool of model.py:190), fused into one kernel:
//   bc[h,w,:] = a0 * (W1 . b1[h/2,w/2]) + a1 * (W2 . b2[h,w])
//             + a2 * max_{2x2}(W3 . b3[2h+i,2w+j]) + a3 * (W4 . max_{4x4} b4[4h+i,4w+j])
// conv-then-pool for b3, pool-then-conv for b4, adds left to right, as the reference.
// One CTA = RF_PIX consecutive output pixels of a row; inputs staged in shared memory.
// ---------------------------------------------------------------------------------
// One CTA = one output row: the stacked 1x1 kernels (K1+K2+K3+K4 rows of N) and the row's inputs are
// staged in shared memory once, then every thread owns (pixel, 4 output channels) items.
__global__ void __launch_bounds__(256)
rfcr_kernel(const float* __restrict__ b1, int ld1, int K1, const float* __restrict__ b2, int ld2, int K2,
            const float* __restrict__ b3, int ld3, int K3, const float* __restrict__ b4, int ld4, int K4,
            const float* __restrict__ wgt, const float* __restrict__ alpha, float* __restrict__ out, int ld_out, int H,
            int W, int N) {
    extern __shared__ __align__(16) float sm[];
    const int KT = K1 + K2 + K3 + K4;
    const int W1 = W >> 1, H1 = H >> 1;
    float* sw = sm;                      // [KT][N]
    float* s1 = sw + (size_t)KT * N;     // [W1][K1]
    float* s2 = s1 + W1 * K1;            // [W][K2]
    float* s3 = s2 + W * K2;             // [W][4][K3]
    float* s4 = s3 + W * 4 * K3;         // [W][K4]   (4x4 max-pooled b4)
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int i = tid; i < KT * N / 4; i += nt) st4(sw + i * 4, ldg4(wgt + (size_t)i * 4));
    for (int i = tid; i < W1 * (K1 / 4); i += nt) {
        const int p = i / (K1 / 4), k = (i % (K1 / 4)) * 4;
        st4(s1 + p * K1 + k, ldg4(b1 + (((size_t)b * H1 + (h >> 1)) * W1 + p) * ld1 + k));
    }
    for (int i = tid; i < W * (K2 / 4); i += nt) {
        const int p = i / (K2 / 4), k = (i % (K2 / 4)) * 4;
        st4(s2 + p * K2 + k, ldg4(b2 + (((size_t)b * H + h) * W + p) * ld2 + k));
    }
    for (int i = tid; i < W * 4 * (K3 / 4); i += nt) {
        const int k = (i % (K3 / 4)) * 4;
        const int q = (i / (K3 / 4)) % 4, p = i / (K3 / 4) / 4;
        st4(s3 + (p * 4 + q) * K3 + k,
            ldg4(b3 + (((size_t)b * 2 * H + 2 * h + (q >> 1)) * (2 * W) + 2 * p + (q & 1)) * ld3 + k));
    }
    for (int i = tid; i < W * (K4 / 4); i += nt) {
        const int p = i / (K4 / 4), k = (i % (K4 / 4)) * 4;
        const float* base = b4 + (((size_t)b * 4 * H + 4 * h) * (4 * W) + 4 * p) * ld4 + k;
        float4 v = ldg4(base);
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (y | x) v = max4(v, ldg4(base + ((size_t)y * 4 * W + x) * ld4));
        st4(s4 + p * K4 + k, v);
    }
    __syncthreads();

    const int N4 = N >> 2;
    const float a0 = __ldg(alpha), a1 = __ldg(alpha + 1), a2 = __ldg(alpha + 2), a3 = __ldg(alpha + 3);
    const float* w1p = sw;
    const float* w2p = w1p + (size_t)K1 * N;
    const float* w3p = w2p + (size_t)K2 * N;
    const float* w4p = w3p + (size_t)K3 * N;
    for (int item = tid; item < W * N4; item += nt) {
        const int p = item / N4, n = (item % N4) * 4;
        auto dot = [&](const float* x, const float* wk, int K) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float xv = x[k];
                const float4 wv = *reinterpret_cast<const float4*>(wk + (size_t)k * N + n);
                a.x = fmaf(xv, wv.x, a.x);
                a.y = fmaf(xv, wv.y, a.y);
                a.z = fmaf(xv, wv.z, a.z);
                a.w = fmaf(xv, wv.w, a.w);
            }
            return a;
        };
        const float4 c1 = dot(s1 + (p >> 1) * K1, w1p, K1);
        const float4 c2 = dot(s2 + p * K2, w2p, K2);
        float4 c3 = dot(s3 + (p * 4 + 0) * K3, w3p, K3);
        c3 = max4(c3, dot(s3 + (p * 4 + 1) * K3, w3p, K3));
        c3 = max4(c3, dot(s3 + (p * 4 + 2) * K3, w3p, K3));
        c3 = max4(c3, dot(s3 + (p * 4 + 3) * K3, w3p, K3));
        const float4 c4 = dot(s4 + p * K4, w4p, K4);
        float4 v;  // ((a0*x0 + a1*x1) + a2*x2) + a3*x3, no FMA contraction across the adds
        v.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, c1.x), __fmul_rn(a1, c2.x)), __fmul_rn(a2, c3.x)), __fmul_rn(a3, c4.x));
        v.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, c1.y), __fmul_rn(a1, c2.y)), __fmul_rn(a2, c3.y)), __fmul_rn(a3, c4.y));
        v.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, c1.z), __fmul_rn(a1, c2.z)), __fmul_rn(a2, c3.z)), __fmul_rn(a3, c4.z));
        v.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, c1.w), __fmul_rn(a1, c2.w)), __fmul_rn(a2, c3.w)), __fmul_rn(a3, c4.w));
        st4(out + (((size_t)b * H + h) * W + p) * ld_out + n, v);
    }
}

int launch_rfcr(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.in2 && op.in3 && op.in4 && op.out && op.w && op.bias, "rfcr: null pointer");
    const int K1 = op.C, K2 = op.K2, K3 = op.K3, K4 = op.K4, N = op.N;
    YR_CHECK_ARG(K1 % 4 == 0 && K2 % 4 == 0 && K3 % 4 == 0 && K4 % 4 == 0 && N % 4 == 0, "rfcr: channels must be multiples of 4");
    YR_CHECK_ARG(op.Ho % 2 == 0 && op.Wo % 2 == 0, "rfcr: output grid must be even");
    YR_CHECK_ARG(op.B <= 65535, "rfcr: batch too large");
    const int W = op.Wo;
    const size_t smem = ((size_t)(K1 + K2 + K3 + K4) * N + (size_t)(W / 2) * K1 + (size_t)W * K2 + (size_t)W * 4 * K3 +
                         (size_t)W * K4) * sizeof(float);
    YR_CHECK_ARG(smem <= 200 * 1024, "rfcr: taps / row too large for shared memory (%zu bytes)", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rfcr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid(op.Ho, op.B);
    rfcr_kernel<<<grid, 256, smem, s>>>((const float*)op.in, op.ld_in, K1, (const float*)op.in2, op.ld_in2, K2,
                                        (const float*)op.in3, op.ld_in3, K3, (const float*)op.in4, op.ld_in4, K4, op.w,
                                        op.bias, (float*)op.out, op.ld_out, op.Ho, op.Wo, N);
    YR_CHECK_LAUNCH("rfcr");
    return YR_OK;
}

// ---------------------------------------------------------------------------------
// Squeeze-excite gate (reference SEBlock, code/yolo3/efficientnet.py:406-438):
//   gate[b,:] = sigmoid(W2^T swish(W1^T mean_hw(x[b]) + b1) + b2)
// One CTA per image; deterministic fixed-order reductions.  The Multiply is folded into
// the A operand of the following project conv (yr_op.scale).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
se_kernel(const float* __restrict__ x, int ld, int HW, int F, int R, const float* __restrict__ w1,
          const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
          float* __restrict__ gate) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int F4 = F >> 2;
    const int P = nt / F4 > 0 ? nt / F4 : 1;  // pixel partitions
    float* part = sm;            // [P][F]
    float* mean = part + P * F;  // [F]
    float* hid = mean + F;       // [R]
    const float* xb = x + (size_t)blockIdx.x * HW * ld;
    for (int c4 = tid % F4, pp = tid / F4; pp < P && c4 < F4; c4 += F4 * P) {  // single pass when nt >= F4
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = pp; p < HW; p += P) {
            const float4 v = ldg4(xb + (size_t)p * ld + c4 * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        st4(part + pp * F + c4 * 4, a);
    }
    __syncthreads();
    for (int f = tid; f < F; f += nt) {
        float a = 0.f;
        for (int pp = 0; pp < P; ++pp) a += part[pp * F + f];
        mean[f] = a / (float)HW;
    }
    __syncthreads();
    for (int r = tid; r < R; r += nt) {
        float a = 0.f;
        for (int f = 0; f < F; ++f) a = fmaf(mean[f], __ldg(w1 + (size_t)f * R + r), a);
        a += __ldg(b1 + r);
        hid[r] = a * (1.0f / (1.0f + expf(-a)));
    }
    __syncthreads();
    for (int f = tid; f < F; f += nt) {
        float a = 0.f;
        for (int r = 0; r < R; ++r) a = fmaf(hid[r], __ldg(w2 + (size_t)r * F + f), a);
        a += __ldg(b2 + f);
        gate[(size_t)blockIdx.x * F + f] = 1.0f / (1.0f + expf(-a));
    }
}

int launch_se(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "se: null pointer");
    const int F = op.C, R = op.N;
    YR_CHECK_ARG(F % 4 == 0 && F / 4 <= 512 && R > 0, "se: unsupported F=%d R=%d", F, R);
    const int nt = 512;
    const int P = nt / (F / 4) > 0 ? nt / (F / 4) : 1;
    const size_t smem = (size_t)(P * F + F + R) * sizeof(float);
    YR_CHECK_ARG(smem <= 48 * 1024, "se: F too large");
    se_kernel<<<op.B, nt, smem, s>>>((const float*)op.in, op.ld_in, op.H * op.W, F, R, op.w, op.bias,
                                     op.w + (size_t)F * R, op.bias + R, (float*)op.out);
    YR_CHECK_LAUNCH("se");
    return YR_OK;
}

// ---------------------------------------------------------------------------------
// Squeeze-excite gate from the per-CTA channel sums the depthwise kernel left behind
// (dwconv.cu): the global mean never re-reads the activation.  One CTA per image.
//   w = [w1t R x F | w2 R x F],  bias = [b1 R | b2 F]   (w1 TRANSPOSED so a warp reads it coalesced)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
se_fc_kernel(const float* __restrict__ part, int slots, int HW, int F, int R, const float* __restrict__ w1t,
             const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
             float* __restrict__ gate) {
    extern __shared__ __align__(16) float sm[];
    float* mean = sm;      // [F]
    float* hid = sm + F;   // [R]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pb = part + (size_t)blockIdx.x * slots * F;
    for (int f4 = tid; f4 < (F >> 2); f4 += 256) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sl = 0; sl < slots; ++sl) {  // fixed order: deterministic
            const float4 v = ldg4(pb + (size_t)sl * F + f4 * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const float d = (float)HW;
        st4(mean + f4 * 4, make_float4(a.x / d, a.y / d, a.z / d, a.w / d));
    }
    __syncthreads();
    for (int r = warp; r < R; r += 8) {
        float a = 0.f;
        for (int f = lane; f < F; f += 32) a = fmaf(mean[f], __ldg(w1t + (size_t)r * F + f), a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            a += __ldg(b1 + r);
            hid[r] = a * (1.0f / (1.0f + expf(-a)));
        }
    }
    __syncthreads();
    // second FC: this CTA's slice of the F outputs (gridDim.y CTAs per image share the work; the squeeze
    // and the first FC above are recomputed per slice - they are tiny, the weight read of FC2 is not)
    const int per = (F + gridDim.y - 1) / gridDim.y;
    const int f_end = min(F, (int)(blockIdx.y + 1) * per);
    for (int f = blockIdx.y * per + tid; f < f_end; f += 256) {
        float a = 0.f;
#pragma unroll 8
        for (int r = 0; r < R; ++r) a = fmaf(hid[r], __ldg(w2 + (size_t)r * F + f), a);
        a += __ldg(b2 + f);
        gate[(size_t)blockIdx.x * F + f] = 1.0f / (1.0f + expf(-a));
    }
}

int launch_se_fc(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "se_fc: null pointer");
    const int F = op.C, R = op.N, slots = op.K2;
    YR_CHECK_ARG(F % 4 == 0 && R > 0 && slots > 0 && op.H > 0 && op.W > 0, "se_fc: unsupported F=%d R=%d slots=%d", F, R, slots);
    const size_t smem = (size_t)(F + R) * sizeof(float);
    YR_CHECK_ARG(smem <= 48 * 1024, "se_fc: F too large");
    YR_CHECK_ARG(op.B <= 65535, "se_fc: batch too large");
    dim3 grid(op.B, F >= 256 ? 4 : (F >= 128 ? 2 : 1));
    se_fc_kernel<<<grid, 256, smem, s>>>((const float*)op.in, slots, op.H * op.W, F, R, op.w, op.bias,
                                         op.w + (size_t)F * R, op.bias + R, (float*)op.out);
    YR_CHECK_LAUNCH("se_fc");
    return YR_OK;
}

}  // namespace yr
