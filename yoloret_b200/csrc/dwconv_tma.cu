// Depthwise 3x3 convolution (stride 1 / 2), NHWC fp32, folded BN + activation: the TMA-staged kernel.
//
// Replaces DepthwiseConv2D+BatchNormalization+ReLU6/Swish of the Keras MobileNetV2 blocks (reference
// code/yolo3/override.py:339), MBConvBlock (code/yolo3/efficientnet.py:501-510) and, with `aux`, the
// squeeze of the SE block that follows it (efficientnet.py:391-403,419).  Per-channel work at 1.7 flop/B:
// no tensor cores, the job is to stream HBM.
//
// A persistent CTA walks (channel tile, spatial tile, image) work items.  For each item ONE
// cp.async.bulk.tensor.4d (TMA, tensor map over [B][H][W][C]) brings the input box
// (rows x cols x 32 channels, halo included) into one of two shared-memory stages; coordinates outside
// the image - TF 'SAME' padding, asymmetric (0,1) for stride 2 on even sizes - and channels >= C are
// zero-filled by the TMA unit, so the compute loop has no bounds checks on its loads and the next item's box
// is in flight while the current one is being computed.  A thread owns 4 channels (one 128-bit lane) of a
// 2 x 4 output patch and walks its input rows once from shared memory (8 lanes cover the 128-byte pixel:
// conflict-free LDS.128); stores are 128-bit, 8 lanes per pixel = full 128-byte lines.
#include "tc_common.cuh"

namespace yr {
namespace dwt {
using namespace yr::tc;

constexpr int CB = 32;        // channels per box (128 bytes)
constexpr int MAX_THREADS = 512;
constexpr int DEFAULT_THREADS = 256;  // measured on B200: 256-thread tiles (two resident CTAs per SM) beat 512 and 128
constexpr int STAGE_LIMIT = 72 * 1024;

struct Params {
    const float* w;
    const float* bias;
    float* out;
    float* part;
    int C, Ho, Wo, ld_out, B;
    int TH, TW;            // output tile of a CTA
    int IH, IW;            // input box
    int cb, cq;            // channels per box (<= 32), float4 lanes per pixel = cb / 4
    int tiles_h, tiles_w, ctiles, total;
    int pad_t, pad_l;
    int stage_floats;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int S, int ACT>
__global__ void __launch_bounds__(MAX_THREADS, 1)
dw_tma_kernel(const __grid_constant__ CUtensorMap tm, const Params p) {
    constexpr int KS = 3, PH = 2, PW = 4;                 // per-thread output patch
    constexpr int IN_ROWS = (PH - 1) * S + KS, IN_COLS = (PW - 1) * S + KS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    float* stage0 = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ float4 s_red[MAX_THREADS / 32][8];
    const uint32_t bar0 = smem_u32(&bars[0]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    }
    __syncthreads();
    pdl_wait();  // programmatic dependent launch: the input is the previous kernel's output
    pdl_launch_dependents();

    const uint32_t stage_bytes = (uint32_t)p.stage_floats * 4u;
    const uint32_t box_bytes = (uint32_t)p.IH * p.IW * p.cb * 4u;
    auto issue = [&](int item, int st) {  // thread 0 only
        const int ct = item % p.ctiles;
        int r = item / p.ctiles;
        const int tw = r % p.tiles_w;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int b = r / p.tiles_h;
        mbar_expect_tx(bar0 + 8u * st, box_bytes);
        tma_load_4d(base + st * stage_bytes, &tm, bar0 + 8u * st, ct * p.cb, tw * p.TW * S - p.pad_l, th * p.TH * S - p.pad_t, b);
    };

    // thread -> (channel quad, patch) of the tile
    const int cq = tid % p.cq;
    const int pg = tid / p.cq;
    const int gw = p.TW / PW;
    const int gx = pg % gw, gy = pg / gw;
    const bool worker = gy < p.TH / PH;

    int item = blockIdx.x;
    if (tid == 0 && item < p.total) issue(item, 0);
    for (uint32_t k = 0; item < p.total; ++k, item += gridDim.x) {
        const int st = k & 1;
        if (tid == 0 && item + (int)gridDim.x < p.total) issue(item + gridDim.x, st ^ 1);
        const int ct = item % p.ctiles;
        int r = item / p.ctiles;
        const int tw = r % p.tiles_w;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int b = r / p.tiles_h;
        const int c = ct * p.cb + cq * 4;
        const bool active = worker && c < p.C;
        float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);

        float4 w[KS * KS];
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
#pragma unroll
            for (int i = 0; i < KS * KS; ++i) w[i] = ldg4(p.w + (size_t)i * p.C + c);
            bv = ldg4(p.bias + c);
        }
        mbar_wait(bar0 + 8u * st, (k >> 1) & 1u, 8);
        if (active) {
            const float* sx = stage0 + (size_t)st * p.stage_floats +
                              ((size_t)(gy * PH * S) * p.IW + gx * PW * S) * p.cb + cq * 4;
            float4 acc[PH][PW];
#pragma unroll
            for (int t = 0; t < PH; ++t)
#pragma unroll
                for (int o = 0; o < PW; ++o) acc[t][o] = bv;  // accumulate on top of the folded-BN bias
#pragma unroll
            for (int rr = 0; rr < IN_ROWS; ++rr) {
                float4 x[IN_COLS];
                const float* row = sx + (size_t)rr * p.IW * p.cb;
#pragma unroll
                for (int j = 0; j < IN_COLS; ++j) x[j] = *reinterpret_cast<const float4*>(row + j * p.cb);
#pragma unroll
                for (int t = 0; t < PH; ++t) {
                    const int kh = rr - t * S;
                    if (kh < 0 || kh >= KS) continue;
#pragma unroll
                    for (int kw = 0; kw < KS; ++kw) {
                        const float4 ww = w[kh * KS + kw];
#pragma unroll
                        for (int o = 0; o < PW; ++o) {
                            fma4(acc[t][o], x[o * S + kw], ww);
                        }
                    }
                }
            }
            const int ho0 = th * p.TH + gy * PH, wo0 = tw * p.TW + gx * PW;
#pragma unroll
            for (int t = 0; t < PH; ++t) {
                const int ho = ho0 + t;
                if (ho >= p.Ho) continue;
#pragma unroll
                for (int o = 0; o < PW; ++o) {
                    const int wo = wo0 + o;
                    if (wo >= p.Wo) continue;
                    float4 v;
                    v.x = apply_act<ACT>(acc[t][o].x);
                    v.y = apply_act<ACT>(acc[t][o].y);
                    v.z = apply_act<ACT>(acc[t][o].z);
                    v.w = apply_act<ACT>(acc[t][o].w);
                    st4(p.out + (((size_t)b * p.Ho + ho) * p.Wo + wo) * p.ld_out + c, v);
                    ps.x += v.x; ps.y += v.y; ps.z += v.z; ps.w += v.w;
                }
            }
        }
        if (p.part != nullptr) {
            // Squeeze-excite 'Mean' fused here: per-item channel sums in a FIXED order (lanes of a warp that share
            // a channel quad by xor-shuffle, then the warps in index order), one [C] slot per spatial tile of the
            // image; se_fc_kernel adds the slots in index order.  No atomics: bit-identical run to run.
            // (cq == 8 lanes per pixel here: SE layers have C % 32 == 0, checked by the launcher.)
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
                ps.x += __shfl_xor_sync(0xffffffffu, ps.x, o);
                ps.y += __shfl_xor_sync(0xffffffffu, ps.y, o);
                ps.z += __shfl_xor_sync(0xffffffffu, ps.z, o);
                ps.w += __shfl_xor_sync(0xffffffffu, ps.w, o);
            }
            if (lane < 8) s_red[warp][lane] = ps;
            __syncthreads();
            if (tid < 8 && ct * p.cb + tid * 4 < p.C) {
                float4 a = s_red[0][tid];
                for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) {
                    const float4 v = s_red[wi][tid];
                    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                }
                const int slots = p.tiles_h * p.tiles_w;
                st4(p.part + (((size_t)b * slots + th * p.tiles_w + tw) * p.C + ct * p.cb + tid * 4), a);
            }
        }
        __syncthreads();  // every thread is done with stage st (and s_red) before the next TMA / item reuses it
    }
}

struct Plan {
    int TH, TW, IH, IW, cb, tiles_h, tiles_w, ctiles, threads, stage_floats;
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Tile choice: the widest tile <= ~56 output columns (28 for stride 2) that cuts the row evenly, then as many
// rows as the thread and shared-memory budgets allow, cut evenly too.
static bool make_plan(const yr_op& op, Plan& pl) {
    const int S = op.stride;
    if (op.k != 3 || (S != 1 && S != 2)) return false;
    if (op.C % 8 || op.ld_in % 4 || op.ld_out % 4) return false;
    pl.cb = op.C < CB ? op.C : CB;
    const int cq = pl.cb / 4;
    const int wmax = S == 1 ? 56 : 28;
    const int nW = cdiv(op.Wo, wmax);
    pl.TW = round_up(cdiv(op.Wo, nW), 4);
    pl.IW = (pl.TW - 1) * S + 3;
    static const int thread_cap = [] {  // developer knob for tile-size experiments
        const char* e = getenv("YR_DW_TMA_THREADS");
        const int v = e ? atoi(e) : 0;
        return v >= 32 && v <= MAX_THREADS ? v : DEFAULT_THREADS;
    }();
    int thmax = 2;
    for (int th = 2; th <= 16; th += 2) {
        const long long threads = (long long)(th / 2) * (pl.TW / 4) * cq;
        const long long bytes = (long long)((th - 1) * S + 3) * pl.IW * pl.cb * 4;
        if (threads <= thread_cap && bytes <= STAGE_LIMIT) thmax = th;
    }
    const int nH = cdiv(op.Ho, thmax);
    pl.TH = round_up(cdiv(op.Ho, nH), 2);
    pl.IH = (pl.TH - 1) * S + 3;
    pl.tiles_h = cdiv(op.Ho, pl.TH);
    pl.tiles_w = cdiv(op.Wo, pl.TW);
    pl.ctiles = cdiv(op.C, pl.cb);
    pl.threads = round_up((pl.TH / 2) * (pl.TW / 4) * cq, 32);
    pl.stage_floats = round_up(pl.IH * pl.IW * pl.cb, 32);
    if (pl.threads > MAX_THREADS || pl.IH > 256 || pl.IW > 256) return false;
    if ((long long)pl.stage_floats * 4 > STAGE_LIMIT) return false;
    if (op.aux && op.C % CB) return false;  // the fused squeeze assumes 8 lanes per pixel
    return true;
}

}  // namespace dwt

// Which depthwise ops run on the TMA kernel (a pure function of the op, so yr_dw_se_slots agrees with the launch).
bool dw_uses_tma(const yr_op& op) {
    static const int mode = [] {
        const char* e = getenv("YR_DW_TMA");
        return e ? atoi(e) : 1;
    }();
    if (!mode) return false;
    dwt::Plan pl;
    if (!dwt::make_plan(op, pl)) return false;
    if (((uintptr_t)op.in) % 16 || ((size_t)op.ld_in * 4) % 16) return false;
    return tc::encode_tiled() != nullptr;
}

int dw_tma_se_slots(const yr_op& op) {
    dwt::Plan pl;
    if (!dwt::make_plan(op, pl)) return 0;
    return pl.tiles_h * pl.tiles_w;
}

template <int S, int ACT>
static int launch_dw_tma_cfg(const yr_op& op, const dwt::Plan& pl, const CUtensorMap& tm, cudaStream_t s) {
    dwt::Params p;
    p.w = op.w;
    p.bias = op.bias;
    p.out = (float*)op.out;
    p.part = op.aux;
    p.C = op.C;
    p.Ho = op.Ho;
    p.Wo = op.Wo;
    p.ld_out = op.ld_out;
    p.B = op.B;
    p.TH = pl.TH;
    p.TW = pl.TW;
    p.IH = pl.IH;
    p.IW = pl.IW;
    p.cb = pl.cb;
    p.cq = pl.cb / 4;
    p.tiles_h = pl.tiles_h;
    p.tiles_w = pl.tiles_w;
    p.ctiles = pl.ctiles;
    p.total = pl.ctiles * pl.tiles_h * pl.tiles_w * op.B;
    p.pad_t = op.pad_t;
    p.pad_l = op.pad_l;
    p.stage_floats = pl.stage_floats;
    const size_t smem = (size_t)2 * pl.stage_floats * 4 + 128;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(dwt::dw_tma_kernel<S, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 2 * dwt::STAGE_LIMIT + 128) != cudaSuccess) {
            set_error("dw_tma: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    const int sms = tc::num_sms();
    // resident CTAs per SM: shared memory, threads and the register file (<= 128 registers per thread)
    int per_sm = 1;
    for (int n = 2; n <= 4; ++n)
        if (smem * n + 4096 * n <= 227 * 1024 && pl.threads * n <= 2048 && pl.threads * n * 128 <= 65536) per_sm = n;
    int grid = sms * per_sm;
    if (grid > p.total) grid = p.total;
    if (launch_pdl(dwt::dw_tma_kernel<S, ACT>, dim3(grid), dim3(pl.threads), smem, s, tm, p) != cudaSuccess) {
        set_error("dw_tma: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    return YR_OK;
}

int launch_dw_tma(const yr_op& op, cudaStream_t s) {
    dwt::Plan pl;
    if (!dwt::make_plan(op, pl)) {
        set_error("dw_tma: no tiling for this op");
        return YR_ERR_UNSUPPORTED;
    }
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) {
        set_error("dw_tma: cuTensorMapEncodeTiled is unavailable in this driver");
        return YR_ERR_CUDA;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[4] = {(cuuint64_t)op.C, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)op.ld_in * 4, (cuuint64_t)op.W * op.ld_in * 4,
                                (cuuint64_t)op.H * op.W * op.ld_in * 4};
    const cuuint32_t box[4] = {(cuuint32_t)pl.cb, (cuuint32_t)pl.IW, (cuuint32_t)pl.IH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(op.in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tc::l2_promotion(),
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("dw_tma: cuTensorMapEncodeTiled failed (%d) for C=%d H=%d W=%d ld=%d box %dx%dx%d", (int)cr, op.C, op.H,
                  op.W, op.ld_in, pl.cb, pl.IW, pl.IH);
        return YR_ERR_CUDA;
    }
    const int key = op.stride * 10 + op.act;
    switch (key) {
        case 10 + YR_ACT_NONE: return launch_dw_tma_cfg<1, YR_ACT_NONE>(op, pl, tm, s);
        case 10 + YR_ACT_RELU6: return launch_dw_tma_cfg<1, YR_ACT_RELU6>(op, pl, tm, s);
        case 10 + YR_ACT_SWISH: return launch_dw_tma_cfg<1, YR_ACT_SWISH>(op, pl, tm, s);
        case 20 + YR_ACT_NONE: return launch_dw_tma_cfg<2, YR_ACT_NONE>(op, pl, tm, s);
        case 20 + YR_ACT_RELU6: return launch_dw_tma_cfg<2, YR_ACT_RELU6>(op, pl, tm, s);
        case 20 + YR_ACT_SWISH: return launch_dw_tma_cfg<2, YR_ACT_SWISH>(op, pl, tm, s);
    }
    set_error("dw_tma: unsupported stride %d / activation %d", op.stride, op.act);
    return YR_ERR_UNSUPPORTED;
}

}  // namespace yr
