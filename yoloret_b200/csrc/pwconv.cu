// Pointwise (1x1) convolution as an fp32 SIMT GEMM with fused epilogue.
//
//   out[m, n] = act( sum_k (A[m,k] * gate[img(m),k]) * W[k,n] + bias[n] ) + res[m,n]
//
// Replaces Conv2D(kernel_size=1)+BatchNormalization(+ReLU6/Swish)(+Add)(+SE Multiply)
// of the reference graph (code/yolo3/model.py:98-114,152-155,243-247,263-267,299-318;
// code/yolo3/efficientnet.py:485-496,517-533; Keras MobileNetV2 expand/project convs).
// This is the exact-fp32 variant (variant=1); the tcgen05 3xTF32 variant lives in
// pwconv_tc.cu.  M = B*H*W is huge and K,N <= 1344, so a CTA owns a 128-row strip
// and streams A exactly once when N fits one column tile.
#include "yr_common.cuh"

namespace yr {

constexpr int PW_BK = 32;
constexpr int PW_APAD = 4;

template <int TX, int TY, int GM, int GN, int ACT, bool HAS_RES, bool HAS_SCALE>
__global__ void __launch_bounds__(TX* TY)
pw_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias,
               const float* __restrict__ res, int ldr, const float* __restrict__ scale, int rows_per_img,
               float* __restrict__ Cout, int ldc, int M, int K, int N) {
    constexpr int NT = TX * TY;
    constexpr int BM = TY * 4 * GM;
    constexpr int BN = TX * 4 * GN;
    constexpr int BK = PW_BK;
    constexpr int AS = BK + PW_APAD;
    constexpr int A_F4 = BM * BK / 4;  // float4s per A tile
    constexpr int B_F4 = BK * BN / 4;
    constexpr int A_PER_T = (A_F4 + NT - 1) / NT;
    constexpr int B_PER_T = (B_F4 + NT - 1) / NT;

    extern __shared__ __align__(16) float smem[];
    float* As = smem;                 // [2][BM][AS]
    float* Bs = smem + 2 * BM * AS;   // [2][BK][BN]

    const int tid = threadIdx.x;
    const int tx = tid % TX;
    const int ty = tid / TX;
    // 1-D grid, column tile fastest: CTAs sharing an A strip run back to back (L2 reuse)
    const int n_tiles = (N + BN - 1) / BN;
    const int m0 = (blockIdx.x / n_tiles) * BM;
    const int n0 = (blockIdx.x % n_tiles) * BN;

    float acc[4 * GM][4 * GN];
#pragma unroll
    for (int i = 0; i < 4 * GM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * GN; ++j) acc[i][j] = 0.0f;

    float4 pa[A_PER_T], pb[B_PER_T];

    auto load_tile = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            const int f = tid + i * NT;
            const int r = f / (BK / 4), k4 = f % (BK / 4);
            const int gm = m0 + r, gk = k0 + k4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((A_F4 % NT == 0 || f < A_F4) && gm < M && gk < K) {
                v = ldg4(A + (size_t)gm * lda + gk);
                if (HAS_SCALE) {
                    const float4 s = ldg4(scale + (size_t)(gm / rows_per_img) * K + gk);
                    v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
                }
            }
            pa[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            const int f = tid + i * NT;
            const int kk = f / (BN / 4), n4 = f % (BN / 4);
            const int gk = k0 + kk, gn = n0 + n4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((B_F4 % NT == 0 || f < B_F4) && gk < K && gn < N) v = ldg4(W + (size_t)gk * N + gn);
            pb[i] = v;
        }
    };
    auto store_tile = [&](int buf) {
        float* as = As + buf * BM * AS;
        float* bs = Bs + buf * BK * BN;
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            const int f = tid + i * NT;
            if (A_F4 % NT == 0 || f < A_F4) st4(as + (f / (BK / 4)) * AS + (f % (BK / 4)) * 4, pa[i]);
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            const int f = tid + i * NT;
            if (B_F4 % NT == 0 || f < B_F4) st4(bs + (f / (BN / 4)) * BN + (f % (BN / 4)) * 4, pb[i]);
        }
    };

    const int KT = (K + BK - 1) / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();

    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_tile(kt + 1);
        const float* as = As + buf * BM * AS;
        const float* bs = Bs + buf * BK * BN;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            float4 a[4 * GM];
#pragma unroll
            for (int g = 0; g < GM; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[g * 4 + i] = *reinterpret_cast<const float4*>(as + (g * TY * 4 + ty * 4 + i) * AS + k4 * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 b[GN];
#pragma unroll
                for (int g = 0; g < GN; ++g)
                    b[g] = *reinterpret_cast<const float4*>(bs + (k4 * 4 + q) * BN + g * TX * 4 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4 * GM; ++i) {
                    const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int g = 0; g < GN; ++g) {
                        acc[i][g * 4 + 0] = fmaf(av, b[g].x, acc[i][g * 4 + 0]);
                        acc[i][g * 4 + 1] = fmaf(av, b[g].y, acc[i][g * 4 + 1]);
                        acc[i][g * 4 + 2] = fmaf(av, b[g].z, acc[i][g * 4 + 2]);
                        acc[i][g * 4 + 3] = fmaf(av, b[g].w, acc[i][g * 4 + 3]);
                    }
                }
            }
        }
        if (kt + 1 < KT) store_tile(buf ^ 1);
        __syncthreads();
    }

    // epilogue: bias (folded BN beta) -> activation -> residual add -> store
#pragma unroll
    for (int g = 0; g < GN; ++g) {
        const int gn = n0 + g * TX * 4 + tx * 4;
        if (gn >= N) continue;
        const float4 bv = ldg4(bias + gn);
#pragma unroll
        for (int gm_ = 0; gm_ < GM; ++gm_)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gm = m0 + gm_ * TY * 4 + ty * 4 + i;
                if (gm >= M) continue;
                const float* ac = acc[gm_ * 4 + i] + g * 4;
                float4 v;
                v.x = apply_act<ACT>(ac[0] + bv.x);
                v.y = apply_act<ACT>(ac[1] + bv.y);
                v.z = apply_act<ACT>(ac[2] + bv.z);
                v.w = apply_act<ACT>(ac[3] + bv.w);
                if (HAS_RES) {
                    const float4 r = ldg4(res + (size_t)gm * ldr + gn);
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                st4(Cout + (size_t)gm * ldc + gn, v);
            }
    }
}

template <int TX, int TY, int GM, int GN, int ACT, bool HAS_RES, bool HAS_SCALE>
static int launch_cfg(const yr_op& op, cudaStream_t s) {
    constexpr int BM = TY * 4 * GM, BN = TX * 4 * GN;
    const int M = op.B * op.H * op.W;
    const size_t smem = (size_t)(2 * BM * (PW_BK + PW_APAD) + 2 * PW_BK * BN) * sizeof(float);
    auto kern = pw_simt_kernel<TX, TY, GM, GN, ACT, HAS_RES, HAS_SCALE>;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid((unsigned)(cdiv(op.N, BN) * cdiv(M, BM)));
    kern<<<grid, TX * TY, smem, s>>>((const float*)op.in, op.ld_in, op.w, op.bias, op.res, op.ld_res, op.scale,
                                     op.H * op.W, (float*)op.out, op.ld_out, M, op.C, op.N);
    YR_CHECK_LAUNCH("pw_simt");
    return YR_OK;
}

template <int ACT, bool HAS_RES, bool HAS_SCALE>
static int launch_tiles(const yr_op& op, cudaStream_t s) {
    const int N = op.N;
    // pick the column tile that wastes the fewest padded columns
    auto waste = [&](int bn) { return cdiv(N, bn) * bn; };
    int best = 128, cost = waste(128);
    const int cands[3] = {96, 64, 32};
    for (int c : cands) {
        // prefer wider tiles on ties (A is re-read once per column tile)
        if (waste(c) < cost) { cost = waste(c); best = c; }
    }
    if (best == 32) return launch_cfg<8, 32, 1, 1, ACT, HAS_RES, HAS_SCALE>(op, s);
    if (best == 64) return launch_cfg<16, 16, 2, 1, ACT, HAS_RES, HAS_SCALE>(op, s);
    if (best == 96) return launch_cfg<8, 32, 1, 3, ACT, HAS_RES, HAS_SCALE>(op, s);
    return launch_cfg<16, 16, 2, 2, ACT, HAS_RES, HAS_SCALE>(op, s);
}

template <int ACT>
static int launch_act(const yr_op& op, cudaStream_t s) {
    const bool r = op.res != nullptr, g = op.scale != nullptr;
    if (r && g) return launch_tiles<ACT, true, true>(op, s);
    if (r) return launch_tiles<ACT, true, false>(op, s);
    if (g) return launch_tiles<ACT, false, true>(op, s);
    return launch_tiles<ACT, false, false>(op, s);
}

int launch_pw(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w && op.bias, "pw: null pointer");
    YR_CHECK_ARG(op.C > 0 && op.C % 8 == 0 && op.N > 0 && op.N % 8 == 0, "pw: K=%d N=%d must be multiples of 8", op.C, op.N);
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= op.N && op.ld_out % 4 == 0,
                 "pw: bad ld_in=%d ld_out=%d", op.ld_in, op.ld_out);
    YR_CHECK_ARG(!op.res || (op.ld_res >= op.N && op.ld_res % 4 == 0), "pw: bad ld_res=%d", op.ld_res);
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w | (uintptr_t)op.bias | (uintptr_t)op.res |
                  (uintptr_t)op.scale) % 16 == 0, "pw: pointers must be 16-byte aligned");
    YR_CHECK_ARG((long long)op.B * op.H * op.W < (1ll << 31), "pw: too many rows");
    switch (op.act) {
        case YR_ACT_NONE: return launch_act<YR_ACT_NONE>(op, s);
        case YR_ACT_RELU6: return launch_act<YR_ACT_RELU6>(op, s);
        case YR_ACT_SWISH: return launch_act<YR_ACT_SWISH>(op, s);
    }
    set_error("pw: unknown activation %d", op.act);
    return YR_ERR_INVALID;
}

}  // namespace yr
