// Pointwise (1x1) convolution on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with the
// 3xTF32 split (a = a_hi + a_lo, w = w_hi + w_lo;  a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi,
// fp32 accumulation in TMEM), which keeps the result within a few fp32 ulps of the exact-fp32
// SIMT variant (pwconv.cu) - plain TF32 would miss the 1e-3 parity budget over ~60 layers.
//
//   out[m, n] = act( sum_k (A[m,k] * gate[img(m),k]) * W[k,n] + bias[n] ) + res[m,n]
//
// Replaces Conv2D(kernel_size=1)+BatchNormalization(+ReLU6/Swish)(+Add)(+SE Multiply) of the
// reference graph (code/yolo3/model.py:98-114,152-155,243-247,263-267,299-318;
// code/yolo3/efficientnet.py:485-496,517-533; Keras MobileNetV2 expand/project convs).
//
// Shape of the problem: M = B*H*W is huge (up to 2.8 M rows), K and N are small (16..720), so
// every layer is a tall-skinny GEMM that streams A once from HBM.  One persistent CTA per SM:
//
//   warp 0      producer   TMA (cp.async.bulk.tensor, SWIZZLE_128B) of 128x32 fp32 A tiles into a
//                          ring of stages; bulk copies (cp.async.bulk) of the pre-split,
//                          pre-swizzled weight image.  If the whole weight slice fits in shared
//                          memory it is loaded ONCE per CTA and stays resident for all M tiles.
//   warps 2-5   converters split each raw A tile in place into (hi, lo) TF32 tiles (elementwise, so
//                          the TMA swizzle is preserved) and fold in the SE gate; fence.proxy.async.
//   warp 1      MMA issuer one thread issues 3 tcgen05.mma per 8-wide K step into one of two TMEM
//                          accumulators; tcgen05.commit releases smem stages / signals the epilogue.
//   warps 6-9   epilogue   tcgen05.ld TMEM -> registers -> 32x33 smem transpose -> bias, activation,
//                          residual -> fully coalesced 128-byte row stores (masking the N/M tails).
#include "yr_common.cuh"
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time

namespace yr {
namespace tc {

constexpr int BM = 128;                      // rows per tile = UMMA M
constexpr int BK = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 4;    // 16 KB (raw/hi) ; lo tile has the same size
constexpr int NUM_THREADS = 320;
constexpr int EPI_LD = 36;                   // padded row (floats) of the per-warp transpose buffer
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
constexpr int SMEM_LIMIT = 232448;           // 227 KB per CTA
constexpr int MAX_A_STAGES = 8, MAX_B_SLOTS = 24;

struct Params {
    const float* wp;      // packed weight image, see pack_kernel
    const float* bias;
    const float* res;
    const float* scale;
    float* out;
    int M, K, N, BN, n_tiles, m_tiles, KB;
    int ld_out, ld_res, rows_per_img, act;
    int nA, nB, resident, tmem_cols, acc_stride, items_per_cta, total_items;
    uint32_t idesc;
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the warp until the phase completes (it wakes
    // on the arrival, ~60 cycles), so idle roles do not steal issue slots from working ones
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (and fails the launch loudly) instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int what) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 24)) {
            printf("yoloret_b200 pw_tc: mbarrier wait timed out (role %d, block %d, thread %d)\n", what, blockIdx.x,
                   threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int what) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, what);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// fp32 -> TF32 (10-bit mantissa), round to nearest / ties away == cvt.rna.tf32.f32, but on the integer
// ALU: the conversion instruction runs on the quarter-rate XU pipe and the converter warps do 64 per tile row.
__host__ __device__ __forceinline__ float tf32_rna(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u = (v.u + 0x1000u) & 0xFFFFE000u;
    return v.f;
#endif
}

// ---- epilogue: TMEM -> registers -> smem transpose -> bias / activation / residual -> 128-bit row stores
// One warp owns TMEM lane quarter q (32 tile rows).  Per 32-column chunk: the accumulator row a lane
// holds goes to smem as 8 STS.128 (row stride 36 floats: conflict-free per 8-lane phase), comes back
// as LDS.128 with 8 lanes covering one row, and leaves as STG.128: 4 full 128-byte lines per store.
constexpr int ACC_FULL_IDX = 3 * MAX_A_STAGES + 2 * MAX_B_SLOTS;

template <int ACT, bool HAS_RES>
__device__ __forceinline__ void epilogue_loop(const Params& p, float* stg, uint32_t tmem_base, uint32_t bar0, int item0,
                                              int item1, int q, int lane) {
    const int sub_r = lane >> 3;        // row within a 4-row group
    const int sub_c = (lane & 7) << 2;  // first of this lane's 4 columns within the chunk
    uint32_t it = 0;
    for (int item = item0; item < item1; ++item, ++it) {
        const int nt = item / p.m_tiles, mt = item - nt * p.m_tiles;
        const uint32_t acc = it & 1;
        mbar_wait(bar0 + 8u * (ACC_FULL_IDX + acc), (it >> 1) & 1, 6);
        tc_fence_after();
        const int row0 = mt * BM + q * 32;
        const int ncols = min(p.BN, p.N - nt * p.BN);  // valid columns of this n tile (multiple of 8)
        const int rows = min(32, p.M - row0);          // valid rows of this warp's slab (may be <= 0)
        const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.acc_stride;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            float v[32];
            tmem_ld32(tsrc + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int c = c0 + sub_c;
            if (c < ncols) {
                const int n = nt * p.BN + c;
                const float4 bv = ldg4(p.bias + n);
                float* op = p.out + (size_t)(row0 + sub_r) * p.ld_out + n;
                const float* rp = HAS_RES ? p.res + (size_t)(row0 + sub_r) * p.ld_res + n : nullptr;
                const float* sp = stg + sub_r * EPI_LD + sub_c;
                const size_t ostep = (size_t)4 * p.ld_out, rstep = (size_t)4 * p.ld_res;
#pragma unroll 4
                for (int r = sub_r; r < rows; r += 4) {
                    float4 x = *reinterpret_cast<const float4*>(sp);
                    x.x = apply_act<ACT>(x.x + bv.x);
                    x.y = apply_act<ACT>(x.y + bv.y);
                    x.z = apply_act<ACT>(x.z + bv.z);
                    x.w = apply_act<ACT>(x.w + bv.w);
                    if (HAS_RES) {
                        const float4 rv = ldg4(rp);
                        x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
                        rp += rstep;
                    }
                    st4(op, x);
                    op += ostep;
                    sp += 4 * EPI_LD;
                }
            }
            __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(bar0 + 8u * (ACC_FULL_IDX + 2 + acc));
    }
}

// ---- the GEMM ------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A stages: (hi 16K | lo 16K) x nA][B slots: (hi | lo) x nB][epilogue transpose][barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_slot_bytes = 2u * p.BN * 128u;
    const uint32_t a_off = 0, b_off = a_off + p.nA * 2u * A_TILE_BYTES;
    const uint32_t epi_off = b_off + p.nB * b_slot_bytes;
    const uint32_t bar_off = epi_off + EPI_BYTES;
    // barrier slots (8 bytes each)
    const uint32_t bar0 = base + bar_off;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_conv = [&](int s) { return bar0 + 8u * (MAX_A_STAGES + s); };
    auto a_empty = [&](int s) { return bar0 + 8u * (2 * MAX_A_STAGES + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (3 * MAX_A_STAGES + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (3 * MAX_A_STAGES + MAX_B_SLOTS + s); };
    auto acc_full = [&](int s) { return bar0 + 8u * (3 * MAX_A_STAGES + 2 * MAX_B_SLOTS + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (3 * MAX_A_STAGES + 2 * MAX_B_SLOTS + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8u * (3 * MAX_A_STAGES + 2 * MAX_B_SLOTS + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nA; ++s) {
            mbar_init(a_full(s), 1);
            mbar_init(a_conv(s), 128);
            mbar_init(a_empty(s), 1);
        }
        for (int s = 0; s < p.nB; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc_full(s), 1);
            mbar_init(acc_empty(s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int item0 = blockIdx.x * p.items_per_cta;
    const int item1 = min(item0 + p.items_per_cta, p.total_items);

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            uint32_t aq = 0, bq = 0;
            bool b_loaded = false;
            for (int item = item0; item < item1; ++item) {
                const int nt = item / p.m_tiles, mt = item - nt * p.m_tiles;
                for (int kb = 0; kb < p.KB; ++kb) {
                    if (!(p.resident && b_loaded)) {
                        const uint32_t slot = bq % p.nB;
                        mbar_wait(b_empty(slot), ((bq / p.nB) & 1) ^ 1, 0);
                        mbar_expect_tx(b_full(slot), b_slot_bytes);
                        bulk_load(base + b_off + slot * b_slot_bytes,
                                  reinterpret_cast<const uint8_t*>(p.wp) + ((size_t)nt * p.KB + kb) * b_slot_bytes,
                                  b_slot_bytes, b_full(slot));
                        ++bq;
                    }
                    const uint32_t st = aq % p.nA;
                    mbar_wait(a_empty(st), ((aq / p.nA) & 1) ^ 1, 1);
                    mbar_expect_tx(a_full(st), A_TILE_BYTES);
                    tma_load_2d(base + a_off + st * 2u * A_TILE_BYTES, &tmA, a_full(st), kb * BK, mt * BM);
                    ++aq;
                }
                b_loaded = true;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t aq = 0, bq = 0, it = 0;
            for (int item = item0; item < item1; ++item, ++it) {
                const uint32_t acc = it & 1;
                mbar_wait(acc_empty(acc), ((it >> 1) & 1) ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.acc_stride;
                for (int kb = 0; kb < p.KB; ++kb) {
                    uint32_t slot;
                    if (p.resident) {
                        slot = kb;
                        mbar_wait(b_full(slot), 0, 3);
                    } else {
                        slot = bq % p.nB;
                        mbar_wait(b_full(slot), (bq / p.nB) & 1, 3);
                        ++bq;
                    }
                    const uint32_t st = aq % p.nA;
                    mbar_wait(a_conv(st), (aq / p.nA) & 1, 4);
                    ++aq;
                    tc_fence_after();
                    const uint32_t a_hi = base + a_off + st * 2u * A_TILE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = base + b_off + slot * b_slot_bytes, b_lo = b_hi + p.BN * 128u;
                    const int ksteps = min(BK, p.K - kb * BK + 7) / 8;  // skip all-zero K steps of the tail
                    for (int k8 = 0; k8 < ksteps; ++k8) {
                        const uint64_t dah = make_desc_sw128(a_hi + k8 * 32), dal = make_desc_sw128(a_lo + k8 * 32);
                        const uint64_t dbh = make_desc_sw128(b_hi + k8 * 32), dbl = make_desc_sw128(b_lo + k8 * 32);
                        umma_tf32(d_tmem, dal, dbh, p.idesc, (kb | k8) ? 1u : 0u);  // small terms first
                        umma_tf32(d_tmem, dah, dbl, p.idesc, 1u);
                        umma_tf32(d_tmem, dah, dbh, p.idesc, 1u);
                    }
                    umma_commit(a_empty(st));
                    if (!p.resident) umma_commit(b_empty(slot));
                }
                umma_commit(acc_full(acc));
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ===== converters: raw fp32 tile -> (hi, lo) TF32 tiles, in place; SE gate folded in =====
        const int ct = threadIdx.x - 64;  // 0..127
        uint32_t aq = 0;
        for (int item = item0; item < item1; ++item) {
            const int mt = item % p.m_tiles;
            for (int kb = 0; kb < p.KB; ++kb, ++aq) {
                const uint32_t st = aq % p.nA;
                mbar_wait(a_full(st), (aq / p.nA) & 1, 5);
                float4* hi = reinterpret_cast<float4*>(gbase + a_off + st * 2u * A_TILE_BYTES);
                float4* lo = reinterpret_cast<float4*>(gbase + a_off + st * 2u * A_TILE_BYTES + A_TILE_BYTES);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int i = ct + j * 128;  // 16-byte chunk index inside the tile
                    float4 v = hi[i];
                    if (p.scale != nullptr) {
                        const int r = i >> 3;
                        const int k = kb * BK + (((i & 7) ^ (r & 7)) << 2);  // undo the 128B swizzle
                        const int row = mt * BM + r;
                        if (k < p.K && row < p.M) {
                            const float4 g = ldg4(p.scale + (size_t)(row / p.rows_per_img) * p.K + k);
                            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
                        }
                    }
                    float4 h, l;
                    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y);
                    l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                    hi[i] = h;
                    lo[i] = l;
                }
                fence_proxy_async();
                mbar_arrive(a_conv(st));
            }
        }
    } else {
        // ===== epilogue =====
        float* stg = reinterpret_cast<float*>(gbase + epi_off) + (warp - 6) * 32 * EPI_LD;
        const bool has_res = p.res != nullptr;
        switch (p.act) {
            case YR_ACT_RELU6:
                if (has_res) epilogue_loop<YR_ACT_RELU6, true>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
                else epilogue_loop<YR_ACT_RELU6, false>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
                break;
            case YR_ACT_SWISH:
                if (has_res) epilogue_loop<YR_ACT_SWISH, true>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
                else epilogue_loop<YR_ACT_SWISH, false>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
                break;
            default:
                if (has_res) epilogue_loop<YR_ACT_NONE, true>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
                else epilogue_loop<YR_ACT_NONE, false>(p, stg, tmem_base, bar0, item0, item1, warp & 3, lane);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// ---- weight packing ---------------------------------------------------------------------------
// W [K][N] row-major (BN scale folded)  ->  for every (n tile, k block): [hi tile | lo tile], each
// BN rows (n) x 32 k-floats in the K-major SWIZZLE_128B image the MMA reads, zero padded.
__global__ void pack_kernel(const float* __restrict__ w, int K, int N, int BN, int n_tiles, int KB,
                            float* __restrict__ packed) {
    const long long total = (long long)n_tiles * KB * BN * BK;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int kk = (int)(idx % BK);
    const int r = (int)((idx / BK) % BN);
    const int kb = (int)((idx / ((long long)BK * BN)) % KB);
    const int nt = (int)(idx / ((long long)BK * BN * KB));
    const int n = nt * BN + r, k = kb * BK + kk;
    const float v = (n < N && k < K) ? w[(size_t)k * N + n] : 0.f;
    const float h = tf32_rna(v), l = tf32_rna(v - h);
    const size_t slot_floats = (size_t)2 * BN * BK;
    const size_t off = (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 32 + (size_t)(((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);
    float* slot = packed + ((size_t)nt * KB + kb) * slot_floats;
    slot[off] = h;
    slot[(size_t)BN * BK + off] = l;
}

// ---- host side ----------------------------------------------------------------------------------
struct Tiling {
    int BN, n_tiles, KB, nA, nB, resident, tmem_cols, acc_stride;
    size_t smem;
};

static bool make_tiling(int K, int N, Tiling& t) {
    if (K <= 0 || N <= 0 || K % 4 || N % 4) return false;
    if (N <= 256) {
        t.n_tiles = 1;
        t.BN = (N + 15) / 16 * 16;
    } else {
        t.n_tiles = (N + 255) / 256;
        t.BN = ((N + t.n_tiles - 1) / t.n_tiles + 15) / 16 * 16;
    }
    if (t.BN < 16) t.BN = 16;
    t.KB = (K + BK - 1) / BK;
    const long long slot = 2ll * t.BN * 128;
    const long long fixed = 1024 /*alignment slack*/ + EPI_BYTES + 1024 /*barriers*/;
    const long long avail = SMEM_LIMIT - fixed;
    const long long a_stage = 2ll * A_TILE_BYTES;
    if (t.n_tiles == 1 && t.KB <= MAX_B_SLOTS && t.KB * slot + 3 * a_stage <= avail) {
        t.resident = 1;
        t.nB = t.KB;
    } else {
        t.resident = 0;
        long long nb = (avail / 2) / slot;
        if (nb < 2) nb = 2;
        if (nb > 4) nb = 4;
        if (nb > t.KB) nb = t.KB;
        t.nB = (int)nb;
        if (t.nB >= t.KB && t.n_tiles == 1 && t.KB <= MAX_B_SLOTS) t.resident = 1;
    }
    long long na = (avail - t.nB * slot) / a_stage;
    if (na > MAX_A_STAGES) na = MAX_A_STAGES;
    if (na < 2) return false;
    t.nA = (int)na;
    t.acc_stride = (t.BN + 31) / 32 * 32;  // the epilogue reads TMEM in 32-column chunks
    int cols = 32;
    while (cols < 2 * t.acc_stride) cols *= 2;
    if (cols > 512) return false;
    t.tmem_cols = cols;
    t.smem = (size_t)(fixed + t.nA * a_stage + t.nB * slot);
    return t.smem <= (size_t)SMEM_LIMIT;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace tc

int launch_pw_tc(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w_tc && op.bias, "pw_tc: null pointer (w_tc = yr_pw_tc_pack output)");
    YR_CHECK_ARG(op.C > 0 && op.C % 8 == 0 && op.N > 0 && op.N % 8 == 0, "pw_tc: K=%d N=%d must be multiples of 8", op.C,
                 op.N);
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= op.N && op.ld_out % 4 == 0,
                 "pw_tc: bad ld_in=%d ld_out=%d", op.ld_in, op.ld_out);
    YR_CHECK_ARG(!op.res || (op.ld_res >= op.N && op.ld_res % 4 == 0), "pw_tc: bad ld_res=%d", op.ld_res);
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w_tc | (uintptr_t)op.bias | (uintptr_t)op.res |
                  (uintptr_t)op.scale) % 16 == 0, "pw_tc: pointers must be 16-byte aligned");
    const long long M = (long long)op.B * op.H * op.W;
    YR_CHECK_ARG(M > 0 && M < (1ll << 31) - 256, "pw_tc: bad row count");
    tc::Tiling t;
    if (!tc::make_tiling(op.C, op.N, t)) {
        set_error("pw_tc: no tiling for K=%d N=%d", op.C, op.N);
        return YR_ERR_UNSUPPORTED;
    }
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) {
        set_error("pw_tc: cuTensorMapEncodeTiled is unavailable in this driver");
        return YR_ERR_CUDA;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)op.C, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)op.ld_in * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(op.in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("pw_tc: cuTensorMapEncodeTiled failed (%d) for K=%d M=%lld ld=%d", (int)cr, op.C, M, op.ld_in);
        return YR_ERR_CUDA;
    }
    tc::Params p;
    p.wp = op.w_tc;
    p.bias = op.bias;
    p.res = op.res;
    p.scale = op.scale;
    p.out = (float*)op.out;
    p.M = (int)M;
    p.K = op.C;
    p.N = op.N;
    p.BN = t.BN;
    p.n_tiles = t.n_tiles;
    p.m_tiles = (int)((M + tc::BM - 1) / tc::BM);
    p.KB = t.KB;
    p.ld_out = op.ld_out;
    p.ld_res = op.ld_res;
    p.rows_per_img = op.H * op.W;
    p.act = op.act;
    p.nA = t.nA;
    p.nB = t.nB;
    p.resident = t.resident;
    p.tmem_cols = t.tmem_cols;
    p.acc_stride = t.acc_stride;
    p.total_items = p.n_tiles * p.m_tiles;
    const int sms = tc::num_sms();
    p.items_per_cta = (p.total_items + sms - 1) / sms;
    const int grid = (p.total_items + p.items_per_cta - 1) / p.items_per_cta;
    // instruction descriptor: D=F32, A=B=TF32, both K-major, N=BN, M=128
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(t.BN >> 3) << 17) | ((uint32_t)(tc::BM >> 4) << 24);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tc::pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT) !=
            cudaSuccess) {
            set_error("pw_tc: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    tc::pw_tc_kernel<<<grid, tc::NUM_THREADS, t.smem, s>>>(tm, p);
    YR_CHECK_LAUNCH("pw_tc");
    return YR_OK;
}

}  // namespace yr

using namespace yr;

extern "C" int64_t yr_pw_tc_packed_floats(int K, int N) {
    tc::Tiling t;
    if (!tc::make_tiling(K, N, t)) return 0;
    return (int64_t)t.n_tiles * t.KB * 2 * t.BN * tc::BK;
}

extern "C" int yr_pw_tc_pack(const float* w, int K, int N, float* packed, void* stream) {
    YR_CHECK_ARG(w && packed, "pw_tc_pack: null pointer");
    tc::Tiling t;
    if (!tc::make_tiling(K, N, t)) {
        set_error("pw_tc_pack: no tensor-core tiling for K=%d N=%d", K, N);
        return YR_ERR_UNSUPPORTED;
    }
    YR_CHECK_ARG(((uintptr_t)packed) % 128 == 0, "pw_tc_pack: packed must be 128-byte aligned");
    const long long total = (long long)t.n_tiles * t.KB * t.BN * tc::BK;
    tc::pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, K, N, t.BN, t.n_tiles, t.KB, packed);
    YR_CHECK_LAUNCH("pw_tc_pack");
    return YR_OK;
}
