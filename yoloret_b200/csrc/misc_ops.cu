// Small streaming kernels of the YOLO-ReT graph: stem conv, nearest/max resampling,
// the fused RFCR fusion and the squeeze-excite gate.  All are HBM/launch bound.
#include "yr_common.cuh"
#include <stdlib.h>

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

// v3 (stride 2): the CTA stages the 2*TH+1 input rows its TH output rows need in shared memory with fully
// coalesced 128-bit loads (a row of the NHWC image with C = 3 is one contiguous run), then a thread owns
// 4 consecutive output pixels x CPT channels: per input row it reads its 27 taps as 7 LDS.128 and reuses every
// weight float4 (a shared-memory broadcast) for 4 pixels, so the loop is ~13 FMAs per shared-memory access
// instead of 4 (v2) and the HBM side is pure streaming.  FMA order per output = v2's: (kh, kw, ci) ascending.
constexpr int STEM_ROUNDS = 3;

template <int ACT, bool U8, int CPT>
__global__ void __launch_bounds__(320, 3)  // 3 CTAs per SM: the kernel is latency-bound at 2 (ncu: no pipe above 45 %)
stem_kernel_v3(const void* __restrict__ in_, const float* __restrict__ wgt, const float* __restrict__ bias,
               float* __restrict__ out, int ld_out, int H, int W, int Ho, int Wo, int N, int pad_t, int pad_l, int TH,
               int row_floats) {
    extern __shared__ __align__(16) float sm3[];
    pdl_wait();
    pdl_launch_dependents();
    float* sw = sm3;                       // [27][N]
    float* sx = sm3 + ((27 * N + 3) & ~3);  // [2*TH+1][row_floats]: pad_l zero pixels, the row, zero tail
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = blockIdx.y, ho0 = blockIdx.x * TH;
    const int in_rows = 2 * TH + 1;
    for (int i = tid; i < 27 * N; i += nt) sw[i] = wgt[i];
    // one warp per 32-chunk segment of a row: zero pads / out-of-image rows, copy the rest (16 bytes per lane;
    // fp32 rows that start 16-byte aligned in shared memory go through cp.async, no register round trip)
    const int lead = pad_l * 3, body = W * 3;
    const int q = body >> 2;                 // 16-byte (fp32) or 4-byte (u8) chunks per row; W % 4 == 0
    const int segs = (q + 31) >> 5;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    // The TH output rows are computed in STEM_ROUNDS rounds of RPR rows; the input rows of every round are
    // requested up front, one cp.async group per round, so round j computes while the rows of j+1.. are in flight.
    const int RPR = TH / STEM_ROUNDS;
    auto stage_rows = [&](int r_begin, int r_end) {
        for (int sg = warp + r_begin * segs; sg < r_end * segs; sg += nwarps) {
            const int r = sg / segs, c4 = (sg - r * segs) * 32 + lane;
            const int hi = ho0 * 2 - pad_t + r;
            float* drow = sx + r * row_floats;
            if (hi < 0 || hi >= H) {
                for (int c = (sg - r * segs) * 128 + lane; c < min(row_floats, (sg - r * segs + 1) * 128); c += 32) drow[c] = 0.0f;
                if (sg - r * segs == segs - 1)
                    for (int c = segs * 128 + lane; c < row_floats; c += 32) drow[c] = 0.0f;
                continue;
            }
            if (sg - r * segs == 0) {  // pads of a valid row
                for (int c = lane; c < lead; c += 32) drow[c] = 0.0f;
                for (int c = lead + body + lane; c < row_floats; c += 32) drow[c] = 0.0f;
            }
            if (c4 >= q) continue;
            if (U8) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(in_) + ((size_t)b * H + hi) * body;
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + c4 * 4));
                float* d = drow + lead + c4 * 4;
                d[0] = (float)(v & 0xffu) * (1.0f / 255.0f);
                d[1] = (float)((v >> 8) & 0xffu) * (1.0f / 255.0f);
                d[2] = (float)((v >> 16) & 0xffu) * (1.0f / 255.0f);
                d[3] = (float)(v >> 24) * (1.0f / 255.0f);
            } else {
                const float* src = reinterpret_cast<const float*>(in_) + ((size_t)b * H + hi) * body + c4 * 4;
                float* d = drow + lead + c4 * 4;
                if (lead == 0) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d)), "l"(src)
                                 : "memory");
                } else {  // lead is a multiple of 3 floats: destination not 16-byte aligned
                    const float4 v = ldg4(src);
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_rows(0, 2 * RPR + 1);
#pragma unroll
    for (int j = 1; j < STEM_ROUNDS; ++j) stage_rows(2 * RPR * j + 1, 2 * RPR * (j + 1) + 1);

    const int NCG = N / CPT;
    const int groups = (Wo + 3) >> 2;
    const int items = RPR * groups * NCG;
#pragma unroll
    for (int round = 0; round < STEM_ROUNDS; ++round) {
    if (round == 0) asm volatile("cp.async.wait_group %0;" ::"n"(STEM_ROUNDS - 1) : "memory");
    else if (round == STEM_ROUNDS - 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.wait_group %0;" ::"n"(STEM_ROUNDS - 2) : "memory");
    __syncthreads();
    for (int item = tid; item < items; item += nt) {
        const int cg = item % NCG;
        const int g = (item / NCG) % groups;
        const int r = round * RPR + item / (NCG * groups);
        const int ho = ho0 + r;
        if (ho >= Ho) break;
        const int n = cg * CPT;
        // channel PAIRS per packed FFMA2 (fma.rn.f32x2: each lane an IEEE fma, so the bits equal scalar fmaf): the
        // kernel is FMA-issue bound (27 taps x CPT channels per pixel), the packed form halves its issue slots
        float2 acc[4][CPT / 2];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int j = 0; j < CPT / 2; ++j) acc[px][j] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            float x[28];
            const float4* xr = reinterpret_cast<const float4*>(sx + (2 * r + kh) * row_floats + g * 24);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const float4 v = xr[i];
                x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float* wp = sw + ((kh * 3 + kw) * 3 + ci) * N + n;
                    float4 wv[CPT / 4];
#pragma unroll
                    for (int j4 = 0; j4 < CPT / 4; ++j4) wv[j4] = *reinterpret_cast<const float4*>(wp + j4 * 4);
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float xv = x[(2 * px + kw) * 3 + ci];
                        const float2 xx = make_float2(xv, xv);
#pragma unroll
                        for (int j4 = 0; j4 < CPT / 4; ++j4) {
                            fma2(acc[px][j4 * 2 + 0], xx, make_float2(wv[j4].x, wv[j4].y));
                            fma2(acc[px][j4 * 2 + 1], xx, make_float2(wv[j4].z, wv[j4].w));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const int wo = g * 4 + px;
            if (wo >= Wo) break;
            float* o = out + (((size_t)b * Ho + ho) * Wo + wo) * ld_out + n;
#pragma unroll
            for (int j4 = 0; j4 < CPT / 4; ++j4) {
                const float4 bv = ldg4(bias + n + j4 * 4);
                float4 v;
                v.x = apply_act<ACT>(acc[px][j4 * 2 + 0].x + bv.x);
                v.y = apply_act<ACT>(acc[px][j4 * 2 + 0].y + bv.y);
                v.z = apply_act<ACT>(acc[px][j4 * 2 + 1].x + bv.z);
                v.w = apply_act<ACT>(acc[px][j4 * 2 + 1].y + bv.w);
                st4(o + j4 * 4, v);
            }
        }
    }
    }
}

template <int ACT, bool U8, int CPT>
static bool launch_stem_v3(const yr_op& op, cudaStream_t s) {
    if (op.stride != 2 || op.W % 4 != 0 || op.B > 65535) return false;
    const int row_items = cdiv(op.Wo, 4) * (op.N / CPT);
    if (row_items > 320) return false;
    const int rows_per_round = 320 / row_items > 0 ? 320 / row_items : 1;
    const int threads = (rows_per_round * row_items + 31) / 32 * 32;
    const int TH = rows_per_round * STEM_ROUNDS;
    // a thread reads 28 floats from float offset 24*g of its rows: the row needs 24*(groups-1) + 28 floats
    int row_floats = (op.pad_l + op.W) * 3;
    const int need = 24 * (cdiv(op.Wo, 4) - 1) + 28;
    if (row_floats < need) row_floats = need;
    row_floats = (row_floats + 3) & ~3;
    const size_t smem = ((size_t)((27 * op.N + 3) & ~3) + (size_t)(2 * TH + 1) * row_floats) * sizeof(float);
    if (smem > 160 * 1024) return false;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        cudaFuncSetAttribute(stem_kernel_v3<ACT, U8, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    dim3 grid(cdiv(op.Ho, TH), op.B);
    launch_pdl(stem_kernel_v3<ACT, U8, CPT>, grid, dim3(threads), smem, s, op.in, op.w, op.bias, (float*)op.out, op.ld_out,
               op.H, op.W, op.Ho, op.Wo, op.N, op.pad_t, op.pad_l, TH, row_floats);
    return true;
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
    // v3 with 4 channels per thread: the N/4 lanes of a pixel store one contiguous run (full 32-byte sectors)
    static const int cpt_knob = [] {  // experiment knob: channels per thread of the staged stem kernel (4 or 8)
        const char* e = getenv("YR_STEM_CPT");
        return e ? atoi(e) : 4;
    }();
    if (cpt_knob == 8 && op.N % 8 == 0 && (op.in_is_u8 ? launch_stem_v3<ACT, true, 8>(op, s) : launch_stem_v3<ACT, false, 8>(op, s))) {
        // 8 channels per thread: each input value is packed once for four FFMA2
    } else if (op.in_is_u8 ? launch_stem_v3<ACT, true, 4>(op, s) : launch_stem_v3<ACT, false, 4>(op, s)) {
    } else if (op.N % 12 == 0) {
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
    pdl_wait();
    pdl_launch_dependents();
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
        launch_pdl(resample_kernel<YR_UP2>, dim3(grid), dim3(256), 0, s, in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
    } else if (op.mode == YR_POOL2) {
        YR_CHECK_ARG(op.Ho == op.H / 2 && op.Wo == op.W / 2, "resample pool2: bad output size");
        launch_pdl(resample_kernel<YR_POOL2>, dim3(grid), dim3(256), 0, s, in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
    } else if (op.mode == YR_POOL4) {
        YR_CHECK_ARG(op.Ho == op.H / 4 && op.Wo == op.W / 4, "resample pool4: bad output size");
        launch_pdl(resample_kernel<YR_POOL4>, dim3(grid), dim3(256), 0, s, in, op.ld_in, out, op.ld_out, op.B, op.H, op.W, op.C, op.Ho, op.Wo);
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
// One CTA = RFCR_ROWS output rows of one image.  The stacked 1x1 kernels (K1+K2+K3+K4 rows of N) and the
// 4x4-max-pooled b4 rows are staged in shared memory once per CTA; a thread owns an item of
// (2 horizontally adjacent output pixels) x (8 output channels): the two pixels share their b1 pixel, every
// weight float4 read from shared memory feeds both, and the activations come straight from global memory as
// float4 over k (they are L1/L2 resident: the 6 channel groups of a pixel pair read the same lines).
// Each dot product keeps the single-accumulator, ascending-k FMA order of a plain 1x1 convolution.
constexpr int RFCR_ROWS = 3;

__device__ __forceinline__ void fma8(float (&acc)[8], float x, const float* w) {
    const float4 w0 = *reinterpret_cast<const float4*>(w);
    const float4 w1 = *reinterpret_cast<const float4*>(w + 4);
    acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
    acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
    acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
    acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
}

__global__ void __launch_bounds__(256)
rfcr_kernel(const float* __restrict__ b1, int ld1, int K1, const float* __restrict__ b2, int ld2, int K2,
            const float* __restrict__ b3, int ld3, int K3, const float* __restrict__ b4, int ld4, int K4,
            const float* __restrict__ wgt, const float* __restrict__ alpha, float* __restrict__ out, int ld_out, int H,
            int W, int N) {
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_launch_dependents();
    const int KT = K1 + K2 + K3 + K4;
    const int W1 = W >> 1, H1 = H >> 1;
    float* sw = sm;                     // [KT][N]
    float* s4 = sw + (size_t)KT * N;    // [RFCR_ROWS][W][K4]   (4x4 max-pooled b4)
    const int h0 = blockIdx.x * RFCR_ROWS, b = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int rows = min(RFCR_ROWS, H - h0);

    for (int i = tid; i < KT * N / 4; i += nt) st4(sw + i * 4, ldg4(wgt + (size_t)i * 4));
    for (int i = tid; i < rows * W * (K4 / 4); i += nt) {
        const int k = (i % (K4 / 4)) * 4;
        const int p = (i / (K4 / 4)) % W, r = i / (K4 / 4) / W;
        const float* base = b4 + (((size_t)b * 4 * H + 4 * (h0 + r)) * (4 * W) + 4 * p) * ld4 + k;
        float4 v = ldg4(base);
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (y | x) v = max4(v, ldg4(base + ((size_t)y * 4 * W + x) * ld4));
        st4(s4 + ((size_t)r * W + p) * K4 + k, v);
    }
    __syncthreads();

    const int NG = N >> 3;
    const float a0 = __ldg(alpha), a1 = __ldg(alpha + 1), a2 = __ldg(alpha + 2), a3 = __ldg(alpha + 3);
    const float* w1p = sw;
    const float* w2p = w1p + (size_t)K1 * N;
    const float* w3p = w2p + (size_t)K2 * N;
    const float* w4p = w3p + (size_t)K3 * N;
    for (int item = tid; item < rows * W1 * NG; item += nt) {
        const int n = (item % NG) * 8;
        const int pp = (item / NG) % W1, r = item / NG / W1;
        const int h = h0 + r, p0 = 2 * pp;

        float c1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // W1 . b1[h/2, w/2]: shared by both pixels
        {
            const float* x = b1 + (((size_t)b * H1 + (h >> 1)) * W1 + pp) * ld1;
#pragma unroll 4
            for (int k = 0; k < K1; k += 4) {  // unrolled: the loads of 4 steps are in flight together (latency-bound loop)
                const float4 xv = ldg4(x + k);
                const float* w = w1p + (size_t)k * N + n;
                fma8(c1, xv.x, w); fma8(c1, xv.y, w + N); fma8(c1, xv.z, w + 2 * N); fma8(c1, xv.w, w + 3 * N);
            }
        }
        float c2[2][8], c3[2][8], c4[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) c2[j][i] = c4[j][i] = 0.f;
        {
            const float* x = b2 + (((size_t)b * H + h) * W + p0) * ld2;
#pragma unroll 2
            for (int k = 0; k < K2; k += 4) {
                const float4 xa = ldg4(x + k), xb = ldg4(x + ld2 + k);
                const float* w = w2p + (size_t)k * N + n;
                fma8(c2[0], xa.x, w); fma8(c2[0], xa.y, w + N); fma8(c2[0], xa.z, w + 2 * N); fma8(c2[0], xa.w, w + 3 * N);
                fma8(c2[1], xb.x, w); fma8(c2[1], xb.y, w + N); fma8(c2[1], xb.z, w + 2 * N); fma8(c2[1], xb.w, w + 3 * N);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {   // max_{2x2}(W3 . b3): conv-then-pool, sub-pixel order (0,0),(0,1),(1,0),(1,1)
            float d[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) d[q][i] = 0.f;
            const float* x = b3 + (((size_t)b * 2 * H + 2 * h) * (2 * W) + 2 * (p0 + j)) * ld3;
            const size_t rs = (size_t)2 * W * ld3;
#pragma unroll 2
            for (int k = 0; k < K3; k += 4) {
                const float4 x0 = ldg4(x + k), x1 = ldg4(x + ld3 + k), x2 = ldg4(x + rs + k), x3 = ldg4(x + rs + ld3 + k);
                const float* w = w3p + (size_t)k * N + n;
                fma8(d[0], x0.x, w); fma8(d[0], x0.y, w + N); fma8(d[0], x0.z, w + 2 * N); fma8(d[0], x0.w, w + 3 * N);
                fma8(d[1], x1.x, w); fma8(d[1], x1.y, w + N); fma8(d[1], x1.z, w + 2 * N); fma8(d[1], x1.w, w + 3 * N);
                fma8(d[2], x2.x, w); fma8(d[2], x2.y, w + N); fma8(d[2], x2.z, w + 2 * N); fma8(d[2], x2.w, w + 3 * N);
                fma8(d[3], x3.x, w); fma8(d[3], x3.y, w + N); fma8(d[3], x3.z, w + 2 * N); fma8(d[3], x3.w, w + 3 * N);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) c3[j][i] = fmaxf(fmaxf(fmaxf(d[0][i], d[1][i]), d[2][i]), d[3][i]);
        }
        {
            const float* x = s4 + ((size_t)r * W + p0) * K4;
            for (int k = 0; k < K4; k += 4) {
                const float4 xa = *reinterpret_cast<const float4*>(x + k), xb = *reinterpret_cast<const float4*>(x + K4 + k);
                const float* w = w4p + (size_t)k * N + n;
                fma8(c4[0], xa.x, w); fma8(c4[0], xa.y, w + N); fma8(c4[0], xa.z, w + 2 * N); fma8(c4[0], xa.w, w + 3 * N);
                fma8(c4[1], xb.x, w); fma8(c4[1], xb.y, w + N); fma8(c4[1], xb.z, w + 2 * N); fma8(c4[1], xb.w, w + 3 * N);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v[8];  // ((a0*x0 + a1*x1) + a2*x2) + a3*x3, no FMA contraction across the adds (WeightedSum, model.py:117-137)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, c1[i]), __fmul_rn(a1, c2[j][i])), __fmul_rn(a2, c3[j][i])),
                                 __fmul_rn(a3, c4[j][i]));
            float* o = out + (((size_t)b * H + h) * W + p0 + j) * ld_out + n;
            st4(o, make_float4(v[0], v[1], v[2], v[3]));
            st4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
}

int launch_rfcr(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.in2 && op.in3 && op.in4 && op.out && op.w && op.bias, "rfcr: null pointer");
    const int K1 = op.C, K2 = op.K2, K3 = op.K3, K4 = op.K4, N = op.N;
    YR_CHECK_ARG(K1 % 4 == 0 && K2 % 4 == 0 && K3 % 4 == 0 && K4 % 4 == 0 && N % 8 == 0, "rfcr: K must be multiples of 4, N of 8");
    YR_CHECK_ARG(op.ld_in % 4 == 0 && op.ld_in2 % 4 == 0 && op.ld_in3 % 4 == 0 && op.ld_in4 % 4 == 0 && op.ld_out % 4 == 0,
                 "rfcr: row strides must be multiples of 4");
    YR_CHECK_ARG(op.Ho % 2 == 0 && op.Wo % 2 == 0, "rfcr: output grid must be even");
    YR_CHECK_ARG(op.B <= 65535, "rfcr: batch too large");
    const int W = op.Wo;
    const size_t smem = ((size_t)(K1 + K2 + K3 + K4) * N + (size_t)RFCR_ROWS * W * K4) * sizeof(float);
    YR_CHECK_ARG(smem <= 200 * 1024, "rfcr: taps / row too large for shared memory (%zu bytes)", smem);
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        cudaFuncSetAttribute(rfcr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid(cdiv(op.Ho, RFCR_ROWS), op.B);
    launch_pdl(rfcr_kernel, grid, dim3(256), smem, s, (const float*)op.in, op.ld_in, K1, (const float*)op.in2, op.ld_in2, K2,
               (const float*)op.in3, op.ld_in3, K3, (const float*)op.in4, op.ld_in4, K4, op.w, op.bias, (float*)op.out,
               op.ld_out, op.Ho, op.Wo, N);
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
// A CLUSTER of SE_CS CTAs serves SE_IPC images: every weight row fetched from L2 feeds SE_IPC dot products, and the
// two FC layers - whose loops are pure L2 latency - are cut SE_CS ways so the chains are short: CTA `rank` computes
// hidden units [rank * R/CS, ...) one row per warp, writes them into EVERY CTA's shared memory through distributed
// shared memory (st.shared::cluster), and after one cluster barrier computes gate channels [rank * F/CS, ...), eight
// threads per channel over disjoint r ranges combined by a fixed xor-shuffle tree.  (The round-1 kernel - one CTA per
// 4 images, 16 CTAs at batch 64 - took 33 us for F = 512: 8 sequential rows per warp, then 128 sequential loads per
// thread.)  Per image the arithmetic and its order do not depend on the batch position or the cluster rank layout.
constexpr int SE_IPC = 4;
constexpr int SE_CS = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(const float* local_smem, uint32_t cta, float v) {
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local_smem), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(512)
se_fc_kernel(const float* __restrict__ part, int slots, int HW, int F, int R, int B, const float* __restrict__ w1t,
             const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
             float* __restrict__ gate) {
    extern __shared__ __align__(16) float sm[];
    // Distributed shared memory may only be written once the target CTA has started executing: arrive now, wait right
    // before the first remote store (compute-sanitizer racecheck flags the stores otherwise).
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    pdl_wait();
    pdl_launch_dependents();
    float* mean = sm;                 // [SE_IPC][F]
    float* hid = sm + SE_IPC * F;     // [SE_IPC][R]  (filled by all CTAs of the cluster)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rank = (int)cluster_ctarank();
    const int img0 = (blockIdx.x / SE_CS) * SE_IPC;
    const int F4 = F >> 2;
    for (int i = tid; i < SE_IPC * F4; i += blockDim.x) {
        const int im = i / F4, f4 = i - im * F4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (img0 + im < B) {
            const float* pb = part + (size_t)(img0 + im) * slots * F;
#pragma unroll 8   // independent loads in flight; the adds keep their fixed order (deterministic)
            for (int sl = 0; sl < slots; ++sl) {
                const float4 v = ldg4(pb + (size_t)sl * F + f4 * 4);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
        }
        const float d = (float)HW;
        st4(mean + im * F + f4 * 4, make_float4(a.x / d, a.y / d, a.z / d, a.w / d));
    }
    __syncthreads();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA of the cluster is running
    // ---- first FC + swish: this CTA's slice of the hidden units, broadcast to the whole cluster ----
    const int RPC = (R + SE_CS - 1) / SE_CS;
    for (int r = rank * RPC + warp; r < min(R, (rank + 1) * RPC); r += nwarps) {
        float a[SE_IPC];
#pragma unroll
        for (int im = 0; im < SE_IPC; ++im) a[im] = 0.f;
#pragma unroll 8   // 8 independent weight loads in flight per lane
        for (int f = lane; f < F; f += 32) {
            const float w = __ldg(w1t + (size_t)r * F + f);
#pragma unroll
            for (int im = 0; im < SE_IPC; ++im) a[im] = fmaf(mean[im * F + f], w, a[im]);
        }
#pragma unroll
        for (int im = 0; im < SE_IPC; ++im) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[im] += __shfl_xor_sync(0xffffffffu, a[im], o);
        }
        const float bb = __ldg(b1 + r);
        if (lane < SE_CS) {  // lane c delivers the row's SE_IPC values to CTA c
#pragma unroll
            for (int im = 0; im < SE_IPC; ++im) {
                const float v = a[im] + bb;
                st_cluster_f32(hid + im * R + r, (uint32_t)lane, v * (1.0f / (1.0f + expf(-v))));
            }
        }
    }
    cluster_sync_all();
    // ---- second FC + sigmoid: this CTA's slice of the gate channels, 8 threads per channel ----
    const int FPC = (F + SE_CS - 1) / SE_CS;
    const int sub = tid & 7;
    const int r_lo = (R * sub) / 8, r_hi = (R * (sub + 1)) / 8;
    for (int fo = tid >> 3; fo < FPC; fo += blockDim.x >> 3) {
        const int f = rank * FPC + fo;
        const bool ok = f < F;
        float a[SE_IPC];
#pragma unroll
        for (int im = 0; im < SE_IPC; ++im) a[im] = 0.f;
        if (ok) {
#pragma unroll 8
            for (int r = r_lo; r < r_hi; ++r) {
                const float w = __ldg(w2 + (size_t)r * F + f);
#pragma unroll
                for (int im = 0; im < SE_IPC; ++im) a[im] = fmaf(hid[im * R + r], w, a[im]);
            }
        }
#pragma unroll
        for (int im = 0; im < SE_IPC; ++im) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) a[im] += __shfl_xor_sync(0xffffffffu, a[im], o);
        }
        if (ok && sub == 0) {
            const float bb = __ldg(b2 + f);
#pragma unroll
            for (int im = 0; im < SE_IPC; ++im)
                if (img0 + im < B) gate[(size_t)(img0 + im) * F + f] = 1.0f / (1.0f + expf(-(a[im] + bb)));
        }
    }
    // no CTA may exit while another can still write into its shared memory: all writes happened before the barrier above
}

int launch_se_fc(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "se_fc: null pointer");
    const int F = op.C, R = op.N, slots = op.K2;
    YR_CHECK_ARG(F % 4 == 0 && R > 0 && slots > 0 && op.H > 0 && op.W > 0, "se_fc: unsupported F=%d R=%d slots=%d", F, R, slots);
    const size_t smem = (size_t)SE_IPC * (F + R) * sizeof(float);
    YR_CHECK_ARG(smem <= 48 * 1024, "se_fc: F too large");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cdiv(op.B, SE_IPC) * SE_CS);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SE_CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, se_fc_kernel, (const float*)op.in, slots, op.H * op.W, F, R, op.B, op.w, op.bias,
                           op.w + (size_t)F * R, op.bias + R, (float*)op.out) != cudaSuccess) {
        set_error("se_fc: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    return YR_OK;
}

}  // namespace yr
