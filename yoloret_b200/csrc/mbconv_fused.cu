// Fused MobileNetV2 inverted-residual block:  expand 1x1 (+BN+ReLU6)  ->  depthwise 3x3 s{1,2} (+BN+ReLU6)
// ->  project 1x1 (+BN) (+ residual), in ONE kernel.  The 6x-expanded tensor and the depthwise output never
// reach HBM: per 128 (s=1: 8x16) or 32 (s=2: 2x16) output pixels the CTA keeps the input halo, one
// 32-channel slice of the expanded tile and of the depthwise output in shared memory / TMEM.
//
// Replaces, per block, the three launches (pw_tc, dw, pw_tc) that implement
//   block_N_expand / expand_BN / expand_relu, block_N_depthwise / BN / relu, block_N_project / BN / block_N_add
// of tf.keras.applications.MobileNetV2 (reference code/yolo3/override.py:339-341) - and the expand ratio-6
// MBConvBlock without SE (reference code/yolo3/efficientnet.py:467-536).  Arithmetic is the unfused path's,
// op for op: both 1x1 convs are 3xTF32 tcgen05 MMAs over the same K order (see pwconv_tc.cu), the depthwise
// taps accumulate in the same (kh, kw) order as dw_kernel, so fused and unfused results are bit-identical.
//
// Roles (320 threads, one persistent CTA per SM):
//   warp 0     producer: one cp.async.bulk of the packed weights at start; per tile one 4-D TMA
//              (cp.async.bulk.tensor.4d, SWIZZLE_128B, out-of-image pixels zero-filled) of the input halo.
//   warp 1     MMA issuer (converged warp, elect.sync): expand MMAs chunk by chunk into a double-buffered
//              TMEM accumulator, project MMAs accumulating over the chunks into a double-buffered D.
//   warps 2-9  workers (256 threads), in lock step per 32-channel chunk:
//              convert (halo -> TF32 hi/lo) | TMEM -> +bias, ReLU6, zero outside the image -> smem E |
//              depthwise 3x3 from E -> +bias, ReLU6 -> TF32 hi/lo A-operand tile P | final epilogue.
#include "tc_common.cuh"

namespace yr {
namespace mb {
using namespace tc;

constexpr int NUM_WORKERS = 256;
constexpr int NUM_THREADS = 64 + NUM_WORKERS;
constexpr int XROWS = 192;              // halo rows a tile may have (s=1: 10x18 = 180, s=2: 5x33 = 165)
constexpr int XBYTES = XROWS * 128;     // one halo buffer: rows of 32 fp32 (128 B, one swizzle row)
constexpr int PBYTES = 128 * 128;       // one half (hi or lo) of the depthwise-output A tile
constexpr int MAX_CHUNKS = 5;           // Ce <= 160
constexpr int SMEM_LIMIT = 232448;
constexpr int TMEM_COLS = 256;          // E: 2 buffers x 2 M tiles x 32 cols = [0,128); D: 2 x 32 = [128,192)
constexpr int D_COL0 = 128;

enum { X_FULL = 0, X_FREE, XC_READY, XHL_FREE, E_FULL0, E_FULL1, E_FREE0, E_FREE1, P_READY, P_FREE, D_FULL0, D_FULL1,
       D_FREE0, D_FREE1, W_FULL, NUM_BARS };

struct Params {
    const float* blob;
    const float* res;
    float* out;
    int B, H, W, Ho, Wo, Cin, Ce, Cout, CoutP, NC;
    int stride, pad_t, pad_l, ld_out, ld_res;
    int TOH, TOW, IH, IW, HR, n_mt, n_px;
    int tiles_x, tiles_y, total_tiles, tiles_per_cta, ks1;
    uint32_t idesc1, idesc2;
    uint32_t off_w1, off_w2, off_wd, off_b1, off_b2, off_b3, blob_bytes;  // byte offsets inside the blob
    long long* dbg;  // optional timeline of CTA 0 (YR_MBCONV_DEBUG=1), else NULL
};

constexpr int DBG_EV = 64;
__device__ __forceinline__ void dbg_mark(const Params& p, int role, uint32_t idx) {
    if (p.dbg != nullptr && blockIdx.x == 0 && idx < DBG_EV) p.dbg[role * DBG_EV + idx] = clock64();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }


// ---- depthwise phase: E (fp32, swizzled rows of 32 channels) -> P (TF32 hi/lo A tile) ----------------
// Thread = (channel quad c4, output column ox, vertical strip): a strip of VS outputs reuses the loaded input
// rows from registers (s=1: 4 outputs from 6x3 taps).  Compile-time stride / tile sizes keep address math cheap.
template <int S>
__device__ __forceinline__ void dw_phase(const uint8_t* __restrict__ e_s, uint8_t* __restrict__ ph_s, uint8_t* __restrict__ pl_s,
                                         const float* __restrict__ wd_s, const float* __restrict__ b2_s, int CeP, int cc, int Ce,
                                         int tid) {
    constexpr int TOW = 16;
    constexpr int IW = (TOW - 1) * S + 3;
    constexpr int VS = S == 1 ? 4 : 2;           // outputs per thread (vertical strip); s=1: 8 rows = 2 strips, s=2: 2 rows = 1 strip
    constexpr int IN_ROWS = (VS - 1) * S + 3;
    const int c4 = tid & 7, ox = (tid >> 3) & 15, strip = tid >> 7;
    const int c = cc * 32 + c4 * 4;
    if (c >= Ce || (S == 2 && strip > 0)) return;
    float4 wv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wv[k] = *reinterpret_cast<const float4*>(wd_s + k * CeP + c);
    const float4 bv = *reinterpret_cast<const float4*>(b2_s + c);
    float4 acc[VS];
#pragma unroll
    for (int r = 0; r < VS; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int oy_base = strip * VS;
#pragma unroll
    for (int ir = 0; ir < IN_ROWS; ++ir) {
        const int hp0 = (oy_base * S + ir) * IW + ox * S;
        float4 x[3];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int hp = hp0 + kw;
            x[kw] = *reinterpret_cast<const float4*>(e_s + hp * 128 + ((c4 ^ (hp & 7)) << 4));
        }
#pragma unroll
        for (int r = 0; r < VS; ++r) {
            const int kh = ir - r * S;
            if (kh < 0 || kh > 2) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float4 ww = wv[kh * 3 + kw];
                acc[r].x = fmaf(x[kw].x, ww.x, acc[r].x);
                acc[r].y = fmaf(x[kw].y, ww.y, acc[r].y);
                acc[r].z = fmaf(x[kw].z, ww.z, acc[r].z);
                acc[r].w = fmaf(x[kw].w, ww.w, acc[r].w);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < VS; ++r) {
        const int op = (oy_base + r) * TOW + ox;
        float4 o, h, l;
        o.x = relu6f(acc[r].x + bv.x); o.y = relu6f(acc[r].y + bv.y);
        o.z = relu6f(acc[r].z + bv.z); o.w = relu6f(acc[r].w + bv.w);
        h.x = tf32_rna(o.x); h.y = tf32_rna(o.y); h.z = tf32_rna(o.z); h.w = tf32_rna(o.w);
        l.x = tf32_rna(o.x - h.x); l.y = tf32_rna(o.y - h.y); l.z = tf32_rna(o.z - h.z); l.w = tf32_rna(o.w - h.w);
        const uint32_t off = op * 128 + ((c4 ^ (op & 7)) << 4);
        *reinterpret_cast<float4*>(ph_s + off) = h;
        *reinterpret_cast<float4*>(pl_s + off) = l;
    }
}

// ---- final epilogue of a tile: D (TMEM) + bias (+ residual) -> out.  All 8 worker warps: the two warps that
// share a TMEM lane quarter take 16 of the (<= 32) output channels each.
struct TileGeo {
    int b, oy0, ox0;
};

__device__ __forceinline__ void epilogue2(const Params& p, const TileGeo& tg, uint32_t tmem_base, uint32_t db, const float* b3_s,
                                          int warp, int w, int lane) {
    const int q4 = warp & 3, half = w >> 2;  // lane quarter, channel half (0: channels 0..15, 1: 16..31)
    const int op = q4 * 32 + lane;
    const int oy = op / p.TOW, ox = op - oy * p.TOW;
    const int gy = tg.oy0 + oy, gx = tg.ox0 + ox;
    const bool ok = op < p.n_px && gy < p.Ho && gx < p.Wo && half * 16 < p.Cout;
    const size_t pix = ((size_t)tg.b * p.Ho + gy) * p.Wo + gx;
    float4 rv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && p.res != nullptr) {  // residual first: its latency overlaps the TMEM load
        const float* r = p.res + pix * p.ld_res + half * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (half * 16 + j * 4 < p.Cout) rv[j] = ldg4(r + j * 4);
    }
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + D_COL0 + db * 32u + (uint32_t)half * 16u)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (ok) {
        float* o = p.out + pix * p.ld_out + half * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (half * 16 + j * 4 < p.Cout) {
                const float4 bv = *reinterpret_cast<const float4*>(b3_s + half * 16 + j * 4);
                float4 x = make_float4(__uint_as_float(u[4 * j]) + bv.x, __uint_as_float(u[4 * j + 1]) + bv.y,
                                       __uint_as_float(u[4 * j + 2]) + bv.z, __uint_as_float(u[4 * j + 3]) + bv.w);
                x.x += rv[j].x; x.y += rv[j].y; x.z += rv[j].z; x.w += rv[j].w;
                st4(o + j * 4, x);
            }
        }
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
mbconv_kernel(const __grid_constant__ CUtensorMap tmX, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
    // carve (all 1024-byte aligned): raw halo | hi | lo | E | P hi | P lo | weight blob | barriers
    const uint32_t o_raw = 0, o_hi = XBYTES, o_lo = 2 * XBYTES, o_e = 3 * XBYTES;
    const uint32_t o_ph = 4 * XBYTES, o_pl = o_ph + PBYTES, o_blob = o_pl + PBYTES;
    const uint32_t o_bar = o_blob + ((p.blob_bytes + 1023u) & ~1023u);
    const uint32_t bar0 = base + o_bar;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(g + o_bar + 8u * NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar(X_FULL), 1);
        mbar_init(bar(X_FREE), NUM_WORKERS);
        mbar_init(bar(XC_READY), NUM_WORKERS);
        mbar_init(bar(XHL_FREE), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(E_FULL0 + i), 1);
            mbar_init(bar(E_FREE0 + i), NUM_WORKERS);
            mbar_init(bar(D_FULL0 + i), 1);
            mbar_init(bar(D_FREE0 + i), NUM_WORKERS);
        }
        mbar_init(bar(P_READY), NUM_WORKERS);
        mbar_init(bar(P_FREE), 1);
        mbar_init(bar(W_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int tile0 = blockIdx.x * p.tiles_per_cta;
    const int tile1 = min(tile0 + p.tiles_per_cta, p.total_tiles);
    const int tiles_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
            mbar_expect_tx(bar(W_FULL), p.blob_bytes);
            bulk_load(base + o_blob, p.blob, p.blob_bytes, bar(W_FULL));
            uint32_t t = 0;
            for (int tile = tile0; tile < tile1; ++tile, ++t) {
                const int b = tile / tiles_img, r = tile - b * tiles_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                mbar_wait(bar(X_FREE), (t & 1u) ^ 1u, 10);
                mbar_expect_tx(bar(X_FULL), (uint32_t)p.HR * 128u);
                tma_load_4d(base + o_raw, &tmX, bar(X_FULL), 0, tx * p.TOW * p.stride - p.pad_l,
                            ty * p.TOH * p.stride - p.pad_t, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint64_t desc0 = make_desc_sw128(base);
        mbar_wait(bar(W_FULL), 0, 11);
        uint32_t t = 0, q = 0;
        auto issue_expand = [&](uint32_t qq, int chunk) {
            const uint32_t eb = qq & 1u, u = qq >> 1;
            mbar_wait(bar(E_FREE0 + eb), (u & 1u) ^ 1u, 12);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dbh = desc0 + ((o_blob + p.off_w1 + (uint32_t)chunk * 8192u) >> 4);
                const uint64_t dbl = dbh + (4096u >> 4);
                for (int mt = 0; mt < p.n_mt; ++mt) {
                    const uint32_t d = tmem_base + eb * 64u + (uint32_t)mt * 32u;
                    const uint64_t dah = desc0 + ((o_hi + (uint32_t)mt * 16384u) >> 4);
                    const uint64_t dal = desc0 + ((o_lo + (uint32_t)mt * 16384u) >> 4);
                    for (int k8 = 0; k8 < p.ks1; ++k8) {
                        const uint64_t ko = (uint64_t)(k8 * 2);
                        umma_tf32(d, dal + ko, dbh + ko, p.idesc1, k8 ? 1u : 0u);
                        umma_tf32(d, dah + ko, dbl + ko, p.idesc1, 1u);
                        umma_tf32(d, dah + ko, dbh + ko, p.idesc1, 1u);
                    }
                }
                umma_commit(bar(E_FULL0 + eb));
                if (chunk == p.NC - 1) umma_commit(bar(XHL_FREE));  // last reader of this tile's halo
            }
            __syncwarp();
        };
        for (int tile = tile0; tile < tile1; ++tile, ++t) {
            const uint32_t db = t & 1u, v = t >> 1;
            mbar_wait(bar(XC_READY), t & 1u, 13);
            tc_fence_after();
            issue_expand(q, 0);
            if (p.NC > 1) issue_expand(q + 1, 1);
            for (int cc = 0; cc < p.NC; ++cc) {
                mbar_wait(bar(P_READY), (q + cc) & 1u, 14);
                if (cc == 0) mbar_wait(bar(D_FREE0 + db), (v & 1u) ^ 1u, 15);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + D_COL0 + db * 32u;
                    const uint64_t dah = desc0 + (o_ph >> 4), dal = desc0 + (o_pl >> 4);
                    const uint64_t dbh = desc0 + ((o_blob + p.off_w2 + (uint32_t)cc * 2u * p.CoutP * 128u) >> 4);
                    const uint64_t dbl = dbh + ((p.CoutP * 128u) >> 4);
                    const int ks = min(4, (p.Ce - cc * 32) / 8);
                    for (int k8 = 0; k8 < ks; ++k8) {
                        const uint64_t ko = (uint64_t)(k8 * 2);
                        umma_tf32(d, dal + ko, dbh + ko, p.idesc2, (cc | k8) ? 1u : 0u);
                        umma_tf32(d, dah + ko, dbl + ko, p.idesc2, 1u);
                        umma_tf32(d, dah + ko, dbh + ko, p.idesc2, 1u);
                    }
                    umma_commit(bar(P_FREE));
                    if (cc == p.NC - 1) umma_commit(bar(D_FULL0 + db));
                }
                __syncwarp();
                if (cc + 2 < p.NC) issue_expand(q + cc + 2, cc + 2);
            }
            q += p.NC;
        }
    } else {
        // ===== workers =====
        const int tid = threadIdx.x - 64, w = tid >> 5;
        const float* blob_s = reinterpret_cast<const float*>(g + o_blob);
        const float* wd_s = reinterpret_cast<const float*>(g + o_blob + p.off_wd);
        const float* b1_s = reinterpret_cast<const float*>(g + o_blob + p.off_b1);
        const float* b2_s = reinterpret_cast<const float*>(g + o_blob + p.off_b2);
        const float* b3_s = reinterpret_cast<const float*>(g + o_blob + p.off_b3);
        (void)blob_s;
        const int CeP = p.NC * 32;
        const float4* xraw = reinterpret_cast<const float4*>(g + o_raw);
        float4* xhi = reinterpret_cast<float4*>(g + o_hi);
        float4* xlo = reinterpret_cast<float4*>(g + o_lo);
        uint8_t* e_s = g + o_e;
        uint8_t* ph_s = g + o_ph;
        uint8_t* pl_s = g + o_pl;
        mbar_wait(bar(W_FULL), 0, 16);
        uint32_t t = 0, q = 0;
        TileGeo prev{0, 0, 0};
        for (int tile = tile0; tile < tile1; ++tile, ++t) {
            const int b = tile / tiles_img, rr = tile - b * tiles_img;
            const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
            const int oy0 = ty * p.TOH, ox0 = tx * p.TOW;
            const int iy0 = oy0 * p.stride - p.pad_t, ix0 = ox0 * p.stride - p.pad_l;
            // ---- halo -> TF32 (hi, lo)
            mbar_wait(bar(X_FULL), t & 1u, 17);
            mbar_wait(bar(XHL_FREE), (t & 1u) ^ 1u, 18);
            if (tid == 0) dbg_mark(p, 0, t);
            for (int i = tid; i < p.HR * 8; i += NUM_WORKERS) {
                const float4 v = xraw[i];
                float4 h, l;
                h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                xhi[i] = h;
                xlo[i] = l;
            }
            fence_proxy_async();
            mbar_arrive(bar(XC_READY));
            mbar_arrive(bar(X_FREE));
            if (tid == 0) dbg_mark(p, 1, t);
            if (t > 0) {  // previous tile: D + bias (+ residual) -> out, while the MMA warp starts this tile's expand
                const uint32_t tp = t - 1;
                mbar_wait(bar(D_FULL0 + (tp & 1u)), (tp >> 1) & 1u, 21);
                tc_fence_after();
                if (tid == 0) dbg_mark(p, 7, tp);
                epilogue2(p, prev, tmem_base, tp & 1u, b3_s, warp, w, lane);
                tc_fence_before();
                mbar_arrive(bar(D_FREE0 + (tp & 1u)));
            }
            // this thread's expanded-tile row (epilogue 1): is its pixel inside the image?
            const int e_mt = w >> 2, e_q = warp & 3;  // TMEM lane quarter = CTA warp index % 4
            const int hp_e = e_mt * 128 + e_q * 32 + lane;
            bool e_inside = false;
            if (e_mt < p.n_mt && hp_e < p.HR) {
                const int iy = hp_e / p.IW, ix = hp_e - iy * p.IW;
                const int gy = iy0 + iy, gx = ix0 + ix;
                e_inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
            }
            for (int cc = 0; cc < p.NC; ++cc, ++q) {
                const uint32_t eb = q & 1u, u = q >> 1;
                // ---- expand accumulator -> +bias, ReLU6, zero outside the image -> E (128 B per halo pixel, swizzled)
                mbar_wait(bar(E_FULL0 + eb), u & 1u, 19);
                tc_fence_after();
                if (tid == 0) dbg_mark(p, 2, q);
                {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(e_q * 32) << 16) + eb * 64u + (uint32_t)e_mt * 32u, v);
                    if (e_mt < p.n_mt && hp_e < p.HR) {
                        uint8_t* row = e_s + hp_e * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bv = *reinterpret_cast<const float4*>(b1_s + cc * 32 + j * 4);
                            float4 o;
                            o.x = e_inside ? relu6f(v[4 * j + 0] + bv.x) : 0.0f;
                            o.y = e_inside ? relu6f(v[4 * j + 1] + bv.y) : 0.0f;
                            o.z = e_inside ? relu6f(v[4 * j + 2] + bv.z) : 0.0f;
                            o.w = e_inside ? relu6f(v[4 * j + 3] + bv.w) : 0.0f;
                            *reinterpret_cast<float4*>(row + ((j ^ (hp_e & 7)) << 4)) = o;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar(E_FREE0 + eb));
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) dbg_mark(p, 3, q);
                // ---- depthwise 3x3 over E -> +bias, ReLU6 -> TF32 (hi, lo) A tile of the project GEMM
                mbar_wait(bar(P_FREE), (q & 1u) ^ 1u, 20);
                if (tid == 0) dbg_mark(p, 4, q);
                if (p.stride == 1) dw_phase<1>(e_s, ph_s, pl_s, wd_s, b2_s, CeP, cc, p.Ce, tid);
                else dw_phase<2>(e_s, ph_s, pl_s, wd_s, b2_s, CeP, cc, p.Ce, tid);
                if (tid == 0) dbg_mark(p, 5, q);
                fence_proxy_async();
                mbar_arrive(bar(P_READY));
                asm volatile("bar.sync 1, 256;" ::: "memory");  // E may be overwritten by the next chunk now
                if (tid == 0) dbg_mark(p, 6, q);
            }
            prev = TileGeo{b, oy0, ox0};  // its final epilogue runs after the NEXT tile's halo conversion
        }
        if (t > 0) {  // final epilogue of the last tile
            const uint32_t tp = t - 1;
            mbar_wait(bar(D_FULL0 + (tp & 1u)), (tp >> 1) & 1u, 21);
            tc_fence_after();
            epilogue2(p, prev, tmem_base, tp & 1u, b3_s, warp, w, lane);
            tc_fence_before();
            mbar_arrive(bar(D_FREE0 + (tp & 1u)));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ---- weight blob ---------------------------------------------------------------------------------
struct Blob {
    int NC, CoutP, CeP;
    uint32_t off_w1, off_w2, off_wd, off_b1, off_b2, off_b3, bytes;
};

static bool make_blob(int Cin, int Ce, int Cout, Blob& b) {
    if (Cin <= 0 || Cin > 32 || Cin % 8 || Ce <= 0 || Ce % 8 || Ce > 32 * MAX_CHUNKS || Cout <= 0 || Cout % 4 || Cout > 32)
        return false;
    b.NC = (Ce + 31) / 32;
    b.CeP = b.NC * 32;
    b.CoutP = (Cout + 15) / 16 * 16;
    b.off_w1 = 0;
    b.off_w2 = b.off_w1 + (uint32_t)b.NC * 8192u;
    b.off_wd = b.off_w2 + (uint32_t)b.NC * 2u * b.CoutP * 128u;
    b.off_b1 = b.off_wd + 9u * b.CeP * 4u;
    b.off_b2 = b.off_b1 + (uint32_t)b.CeP * 4u;
    b.off_b3 = b.off_b2 + (uint32_t)b.CeP * 4u;
    b.bytes = b.off_b3 + 32u * 4u;
    b.bytes = (b.bytes + 15u) & ~15u;
    return true;
}

__global__ void pack_kernel(const float* __restrict__ w1, int ld1, const float* __restrict__ b1, const float* __restrict__ wd,
                            int ldd, const float* __restrict__ b2, const float* __restrict__ w2, int ld2,
                            const float* __restrict__ b3, int Cin, int Ce, int Cout, Blob bl, float* __restrict__ blob) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_w1 = bl.NC * 32 * 32, n_w2 = bl.NC * bl.CoutP * 32, n_wd = 9 * bl.CeP;
    if (idx < n_w1) {  // W1 [Cin][Ce] -> per chunk: (hi | lo) 32 rows (n) x 32 k, K-major SWIZZLE_128B
        const int k = idx % 32, r = (idx / 32) % 32, c = idx / 1024;
        const int n = c * 32 + r;
        const float v = (n < Ce && k < Cin) ? w1[(size_t)k * ld1 + n] : 0.f;
        const float h = tf32_rna(v), l = tf32_rna(v - h);
        const int off = (r >> 3) * 256 + (r & 7) * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3);
        float* dst = blob + bl.off_w1 / 4 + c * 2048;
        dst[off] = h;
        dst[1024 + off] = l;
    } else if (idx < n_w1 + n_w2) {  // W2 [Ce][Cout] -> per chunk: (hi | lo) CoutP rows (n) x 32 k
        const int i = idx - n_w1;
        const int kk = i % 32, r = (i / 32) % bl.CoutP, c = i / (32 * bl.CoutP);
        const int k = c * 32 + kk;
        const float v = (r < Cout && k < Ce) ? w2[(size_t)k * ld2 + r] : 0.f;
        const float h = tf32_rna(v), l = tf32_rna(v - h);
        const int off = (r >> 3) * 256 + (r & 7) * 32 + (((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);
        float* dst = blob + bl.off_w2 / 4 + c * 2 * bl.CoutP * 32;
        dst[off] = h;
        dst[bl.CoutP * 32 + off] = l;
    } else if (idx < n_w1 + n_w2 + n_wd) {
        const int i = idx - n_w1 - n_w2;
        const int c = i % bl.CeP, k = i / bl.CeP;
        blob[bl.off_wd / 4 + i] = c < Ce ? wd[(size_t)k * ldd + c] : 0.f;
    } else if (idx < n_w1 + n_w2 + n_wd + 2 * bl.CeP + 32) {
        const int i = idx - n_w1 - n_w2 - n_wd;
        if (i < bl.CeP) blob[bl.off_b1 / 4 + i] = i < Ce ? b1[i] : 0.f;
        else if (i < 2 * bl.CeP) blob[bl.off_b2 / 4 + (i - bl.CeP)] = (i - bl.CeP) < Ce ? b2[i - bl.CeP] : 0.f;
        else blob[bl.off_b3 / 4 + (i - 2 * bl.CeP)] = (i - 2 * bl.CeP) < Cout ? b3[i - 2 * bl.CeP] : 0.f;
    }
}

static size_t smem_bytes(const Blob& b) {
    return 1024 + 4 * (size_t)XBYTES + 2 * (size_t)PBYTES + ((b.bytes + 1023u) & ~1023u) + 256;
}

}  // namespace mb

int launch_mbconv(const yr_op& op, cudaStream_t s) {
    using namespace mb;
    YR_CHECK_ARG(op.in && op.out && op.w_tc, "mbconv: null pointer (w_tc = yr_mbconv_pack output)");
    YR_CHECK_ARG(op.k == 3 && (op.stride == 1 || op.stride == 2), "mbconv: depthwise must be 3x3 stride 1/2");
    const int Cin = op.C, Ce = op.K2, Cout = op.N;
    Blob bl;
    if (!make_blob(Cin, Ce, Cout, bl) || smem_bytes(bl) > (size_t)SMEM_LIMIT) {
        set_error("mbconv: unsupported channels Cin=%d Ce=%d Cout=%d", Cin, Ce, Cout);
        return YR_ERR_UNSUPPORTED;
    }
    YR_CHECK_ARG(op.ld_in >= Cin && op.ld_in % 4 == 0 && op.ld_out >= Cout && op.ld_out % 4 == 0, "mbconv: bad ld");
    YR_CHECK_ARG(!op.res || (op.ld_res >= Cout && op.ld_res % 4 == 0), "mbconv: bad ld_res");
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w_tc | (uintptr_t)op.res) % 16 == 0,
                 "mbconv: pointers must be 16-byte aligned");
    YR_CHECK_ARG(op.B > 0 && op.H > 0 && op.W > 0 && op.Ho > 0 && op.Wo > 0, "mbconv: bad sizes");
    Params p;
    p.blob = op.w_tc;
    p.res = op.res;
    p.out = (float*)op.out;
    p.B = op.B; p.H = op.H; p.W = op.W; p.Ho = op.Ho; p.Wo = op.Wo;
    p.Cin = Cin; p.Ce = Ce; p.Cout = Cout; p.CoutP = bl.CoutP; p.NC = bl.NC;
    p.stride = op.stride; p.pad_t = op.pad_t; p.pad_l = op.pad_l; p.ld_out = op.ld_out; p.ld_res = op.ld_res;
    p.TOW = 16;
    p.TOH = op.stride == 1 ? 8 : 2;
    p.IH = (p.TOH - 1) * op.stride + 3;
    p.IW = (p.TOW - 1) * op.stride + 3;
    p.HR = p.IH * p.IW;
    p.n_mt = (p.HR + 127) / 128;
    p.n_px = p.TOH * p.TOW;
    YR_CHECK_ARG(p.HR <= XROWS && p.n_mt <= 2, "mbconv: halo too large");
    p.tiles_x = (op.Wo + p.TOW - 1) / p.TOW;
    p.tiles_y = (op.Ho + p.TOH - 1) / p.TOH;
    const long long total = (long long)op.B * p.tiles_x * p.tiles_y;
    YR_CHECK_ARG(total < (1ll << 31), "mbconv: too many tiles");
    p.total_tiles = (int)total;
    const int sms = num_sms();
    p.tiles_per_cta = (p.total_tiles + sms - 1) / sms;
    const int grid = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    p.ks1 = (Cin + 7) / 8;
    p.idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bl.CoutP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.off_w1 = bl.off_w1; p.off_w2 = bl.off_w2; p.off_wd = bl.off_wd;
    p.off_b1 = bl.off_b1; p.off_b2 = bl.off_b2; p.off_b3 = bl.off_b3; p.blob_bytes = bl.bytes;

    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("mbconv: cuTensorMapEncodeTiled is unavailable in this driver");
        return YR_ERR_CUDA;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)op.ld_in * 4, (cuuint64_t)op.W * op.ld_in * 4,
                                (cuuint64_t)op.H * op.W * op.ld_in * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)p.IW, (cuuint32_t)p.IH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(op.in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("mbconv: cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return YR_ERR_CUDA;
    }
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(mbconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
            set_error("mbconv: cannot raise the dynamic shared memory limit");
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    p.dbg = nullptr;
    static const bool debug = getenv("YR_MBCONV_DEBUG") != nullptr;  // developer aid only: timeline of CTA 0
    if (debug) {
        static long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc(&dbuf, 8 * DBG_EV * sizeof(long long));
        cudaMemsetAsync(dbuf, 0, 8 * DBG_EV * sizeof(long long), s);
        p.dbg = dbuf;
    }
    mbconv_kernel<<<grid, NUM_THREADS, smem_bytes(bl), s>>>(tm, p);
    YR_CHECK_LAUNCH("mbconv");
    if (debug) {
        static long long h[8 * DBG_EV];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[8] = {"x_ready(t)", "conv_done(t)", "e_full(q)", "epi1_done(q)", "p_free(q)", "dw_done(q)",
                                "chunk_end(q)", "d_full(t)"};
        const long long t0 = h[0];
        fprintf(stderr, "mbconv timeline Cin=%d Ce=%d Cout=%d s=%d NC=%d tiles/cta=%d (cycles since first x_ready)\n", Cin, Ce,
                Cout, op.stride, p.NC, p.tiles_per_cta);
        for (int r = 0; r < 8; ++r) {
            fprintf(stderr, "%-14s", names[r]);
            for (int i = 0; i < 22 && h[r * DBG_EV + i]; ++i) fprintf(stderr, " %6lld", h[r * DBG_EV + i] - t0);
            fprintf(stderr, "\n");
        }
    }
    return YR_OK;
}

}  // namespace yr

using namespace yr;

extern "C" int64_t yr_mbconv_packed_floats(int Cin, int Ce, int Cout) {
    mb::Blob b;
    if (!mb::make_blob(Cin, Ce, Cout, b) || mb::smem_bytes(b) > (size_t)mb::SMEM_LIMIT) return 0;
    return (int64_t)b.bytes / 4;
}

extern "C" int yr_mbconv_pack(const float* w1, int ld1, const float* b1, const float* wd, int ldd, const float* b2,
                              const float* w2, int ld2, const float* b3, int Cin, int Ce, int Cout, float* packed,
                              void* stream) {
    YR_CHECK_ARG(w1 && b1 && wd && b2 && w2 && b3 && packed, "mbconv_pack: null pointer");
    mb::Blob b;
    if (!mb::make_blob(Cin, Ce, Cout, b)) {
        set_error("mbconv_pack: unsupported channels Cin=%d Ce=%d Cout=%d", Cin, Ce, Cout);
        return YR_ERR_UNSUPPORTED;
    }
    YR_CHECK_ARG(((uintptr_t)packed) % 128 == 0, "mbconv_pack: packed must be 128-byte aligned");
    const int total = b.NC * 1024 + b.NC * b.CoutP * 32 + 9 * b.CeP + 2 * b.CeP + 32;
    mb::pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, ld1, b1, wd, ldd, b2, w2, ld2, b3, Cin, Ce, Cout,
                                                                            b, packed);
    YR_CHECK_LAUNCH("mbconv_pack");
    return YR_OK;
}
