// Pointwise (1x1) convolution on the 5th-gen tensor cores, A operand in TENSOR MEMORY ("TS" form of
// tcgen05.mma): the second-generation kernel behind YR_PW_TS.  Same arithmetic as pwconv_tc.cu -
// 3xTF32 split, fp32 accumulation in TMEM,
//
//   out[m, n] = act( sum_k (A[m,k] * gate[img(m),k]) * W[k,n] + bias[n] ) + res[m,n]
//
// (Conv2D 1x1 + BatchNormalization + ReLU6/Swish + Add + SE Multiply of the reference graph,
// code/yolo3/model.py:98-114,152-155,243-247,263-267,299-318; code/yolo3/efficientnet.py:485-496,517-533)
// - but the split activation tiles never go back to shared memory.
//
// Why: the ncu/timeline study of pwconv_tc.cu (profiles/README.md) showed its K-heavy layers are bound by
// SHARED-MEMORY BANDWIDTH, not HBM: per 16 KB of A that arrives from HBM the SM moved 16 KB (TMA write)
// + 16 KB (converter read) + 32 KB (converter writes hi and lo) + 48 KB (three MMAs read A) + the weight
// tiles, ~1100 cycles at 128 B/clk against the ~700 the HBM rate needs.  Here the converters read the raw
// tile once (16 KB), split it in registers and write (hi, lo) straight into TMEM with tcgen05.st
// (256 B/clk, a separate datapath); the MMAs read A from TMEM and only the weight tiles from shared
// memory.  The raw-tile slot is released as soon as the converter has read it, and the (hi, lo) ring
// costs no shared memory, so more layers keep their whole weight image resident.
//
//   warp 0       A producer  TMA (SWIZZLE_128B) of 128x32 fp32 A tiles into a ring
//   warp 18      W producer  bulk copies of the pre-split, pre-swizzled weight image (resident when it fits,
//                            else a ring of up to 5 slots: the MMA issuer's wait is L2 latency, so depth matters)
//   warps 2-9    converters  2 groups x 4 warps alternate k-blocks; a thread owns one tile row (= its TMEM
//                            lane): 8 conflict-free LDS.128, SE gate, hi/lo split, 4 tcgen05.st.x16
//   warp 1       MMA issuer  3 tcgen05.mma (A in TMEM, B descriptor) per 8-wide K step; tcgen05.commit
//                            frees the TMEM stage / weight slot and signals the epilogue
//   warps 10-17  epilogue    2 groups x 4 warps alternate tiles: tcgen05.ld -> smem transpose -> bias,
//                            activation, residual -> 128-bit coalesced row stores
//
// TMEM map (512 columns): [nAcc accumulators x acc_stride][nT A stages x (32 hi + 32 lo) columns].
// Output channels are cut into n tiles of <= 192 columns; work items run m-major so the n tiles of one
// 128-row block are consecutive and re-read their A tile from L2.
#include "tc_common.cuh"

namespace yr {
namespace ts {
using namespace yr::tc;

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int NUM_CONVERTERS = 256;
constexpr int NUM_EPILOGUE = 256;
constexpr int NUM_THREADS = 64 + NUM_CONVERTERS + NUM_EPILOGUE + 32;  // + A producer, MMA issuer, weight producer warps
constexpr int EPI_LD = 36;
constexpr int EPI_STAGE_BYTES = 4 * 32 * EPI_LD * 4;  // per epilogue group: 4 transpose buffers (+ the bias copy)
constexpr int SMEM_LIMIT = 232448;
constexpr int MAX_A = 12, MAX_T = 4, MAX_B = 32, MAX_ACC = 4;
constexpr int BAR_BYTES = 1024;
constexpr int T_STAGE_COLS = 64;  // hi columns [0,32), lo columns [32,64)

struct Params {
    const float* wp;
    const float* bias;
    const float* res;
    const float* scale;
    float* out;
    int M, K, N, BN, n_tiles, m_tiles, KB;
    int ld_out, ld_res, rows_per_img, act;
    int up2, img_w;  // up2: every output row is stored to its 2x2 nearest-upsampled pixels (fused UpSampling2D)
    int nA, nT, nB, nAcc, resident, acc_stride, a_col0, items_per_cta, total_items, epi_group_bytes;
    uint32_t idesc;
    long long* dbg;
};

constexpr int BAR_A_FULL = 0;
constexpr int BAR_A_EMPTY = BAR_A_FULL + MAX_A;
constexpr int BAR_T_FULL = BAR_A_EMPTY + MAX_A;
constexpr int BAR_T_EMPTY = BAR_T_FULL + MAX_T;
constexpr int BAR_B_FULL = BAR_T_EMPTY + MAX_T;
constexpr int BAR_B_EMPTY = BAR_B_FULL + MAX_B;
constexpr int BAR_ACC_FULL = BAR_B_EMPTY + MAX_B;
constexpr int BAR_ACC_EMPTY = BAR_ACC_FULL + MAX_ACC;
constexpr int BAR_COUNT = BAR_ACC_EMPTY + MAX_ACC;
static_assert(BAR_COUNT * 8 + 8 <= BAR_BYTES, "barrier table overflows its reservation");

constexpr int DBG_EV = 256;
// The timeline marks are compiled in only for the debug instantiation of the kernel (YR_PW_TC_DEBUG=1): in the MMA
// issuer's serial loop even a predicated-off clock read and store are instructions on the critical path.
template <bool DBG>
__device__ __forceinline__ void dbg_mark_t(const Params& p, int role, uint32_t idx) {
    if (DBG) {
        if (p.dbg != nullptr && blockIdx.x == 0 && idx < DBG_EV) p.dbg[role * DBG_EV + idx] = clock64();
    }
}
#define dbg_mark(p, role, idx) dbg_mark_t<DBG>(p, role, idx)

// tcgen05.mma with the A operand in tensor memory (lane = tile row, one TF32 element per column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// registers -> 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- epilogue (see pwconv_tc.cu for the access pattern; items here are m-major) ---------------
template <int ACT, bool HAS_RES, bool UP2, bool DBG>
__device__ __forceinline__ void epilogue_loop(const Params& p, float* stg, const float* s_bias, uint32_t tmem_base,
                                              uint32_t bar0, int item0, int item1, int q, int lane, int ewarp, int grp) {
    const int sub_r = lane >> 3;
    const int sub_c = (lane & 7) << 2;
    uint32_t it = 0;
    Ring racc;
    int mt = item0 / p.n_tiles, nt = item0 - mt * p.n_tiles;
    for (int item = item0; item < item1; ++item, ++it) {
        if ((int)(it & 1u) == grp) {
            const uint32_t acc = racc.slot;
            const int row0 = mt * BM + q * 32;
            const int ncols = min(p.BN, p.N - nt * p.BN);
            const int rows = min(32, p.M - row0);
            const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.acc_stride;
            bool waited = false;
            float v[32];
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                const int c = c0 + sub_c;
                const int n = nt * p.BN + c;
                const bool col_ok = c < ncols;
                float4 rv[8];
                if (HAS_RES) {
                    const float* rp = p.res + (size_t)(row0 + sub_r) * p.ld_res + n;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        rv[i] = (col_ok && sub_r + 4 * i < rows) ? ldg4(rp + (size_t)(4 * i) * p.ld_res)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (!waited) {
                    mbar_wait(bar0 + 8u * (BAR_ACC_FULL + acc), racc.phase, 6);
                    tc_fence_after();
                    waited = true;
                    if (ewarp == 0 && lane == 0) dbg_mark(p, 6, it);
                    if (!HAS_RES) tmem_ld32_issue(tsrc + c0, v);
                }
                if (HAS_RES) tmem_ld32_issue(tsrc + c0, v);  // residual variant: no prefetch (register budget)
                tmem_ld_wait();  // this chunk's accumulator columns are in v
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * j) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                // v now lives in shared memory: fetch the next chunk into the same registers while this one is stored
                if (!HAS_RES && c0 + 32 < ncols) tmem_ld32_issue(tsrc + c0 + 32, v);
                __syncwarp();
                if (col_ok) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + n);
                    float* op = p.out + (size_t)(row0 + sub_r) * p.ld_out + n;
                    const float* sp = stg + sub_r * EPI_LD + sub_c;
                    const size_t ostep = (size_t)4 * p.ld_out;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (sub_r + 4 * i < rows) {
                            float4 x = *reinterpret_cast<const float4*>(sp + (4 * i) * EPI_LD);
                            x.x = apply_act<ACT>(x.x + bv.x);
                            x.y = apply_act<ACT>(x.y + bv.y);
                            x.z = apply_act<ACT>(x.z + bv.z);
                            x.w = apply_act<ACT>(x.w + bv.w);
                            if (HAS_RES) { x.x += rv[i].x; x.y += rv[i].y; x.z += rv[i].z; x.w += rv[i].w; }
                            if (!UP2) {
                                st4(op + (size_t)i * ostep, x);
                            } else {
                                // fused UpSampling2D (nearest x2, reference code/yolo3/model.py:254,274): input pixel
                                // (b, h, w) -> output pixels (2h..2h+1, 2w..2w+1) of the [B, 2H, 2W, ld_out] tensor
                                const int m = row0 + sub_r + 4 * i;
                                const int bi = m / p.rows_per_img, rem = m - bi * p.rows_per_img;
                                const int h = rem / p.img_w, w = rem - h * p.img_w;
                                const size_t wo2 = (size_t)2 * p.img_w;
                                float* d = p.out + (((size_t)bi * 2 * (p.rows_per_img / p.img_w) + 2 * h) * wo2 + 2 * w) * p.ld_out + n;
                                st4(d, x);
                                st4(d + p.ld_out, x);
                                st4(d + wo2 * p.ld_out, x);
                                st4(d + (wo2 + 1) * p.ld_out, x);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (!waited) {
                mbar_wait(bar0 + 8u * (BAR_ACC_FULL + acc), racc.phase, 6);
                tc_fence_after();
            }
            tc_fence_before();
            mbar_arrive(bar0 + 8u * (BAR_ACC_EMPTY + acc));
            if (ewarp == 0 && lane == 0) dbg_mark(p, 7, it);
        }
        racc.advance(p.nAcc);
        if (++nt == p.n_tiles) { nt = 0; ++mt; }
    }
}

// ---- converters: raw fp32 A tile (smem) -> (hi, lo) TF32 columns of a TMEM stage ----------------
// Thread = tile row r = its TMEM lane.  The 128-byte row is read as 8 LDS.128 whose chunk index is
// XORed with (r & 7) (the TMA swizzle): the 8 lanes of a quarter-warp hit 8 different bank groups.
template <bool HAS_SCALE, bool DBG>
__device__ __forceinline__ void converter_loop(const Params& p, const uint8_t* a_ring, uint32_t tmem_base, uint32_t bar0,
                                               int item0, int item1, int q, int lane, int grp) {
    Ring ra, rt;
    uint32_t dq = 0;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0;
    int mt = item0 / p.n_tiles, nt = item0 - mt * p.n_tiles;
    for (int item = item0; item < item1; ++item) {
        for (int kb = 0; kb < p.KB; ++kb, ++dq) {
            if ((int)(dq & 1u) == grp) {
                mbar_wait(bar0 + 8u * (BAR_A_FULL + ra.slot), ra.phase, 5);
                if (q == 0 && lane == 0) dbg_mark(p, 1, dq);
                const uint8_t* row = a_ring + ra.slot * (uint32_t)A_TILE_BYTES + r * 128;
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(row + ((c ^ (r & 7)) << 4));
                if (HAS_SCALE) {
                    const int grow = mt * BM + r;
                    if (grow < p.M) {
                        const float* g = p.scale + (size_t)(grow / p.rows_per_img) * p.K + kb * BK;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (kb * BK + 4 * c < p.K) {
                                const float4 gv = ldg4(g + 4 * c);
                                v[c].x *= gv.x; v[c].y *= gv.y; v[c].z *= gv.z; v[c].w *= gv.w;
                            }
                        }
                    }
                }
                mbar_wait(bar0 + 8u * (BAR_T_EMPTY + rt.slot), rt.phase ^ 1u, 7);
                tc_fence_after();
                if (q == 0 && lane == 0) dbg_mark(p, 2, dq);
                const uint32_t taddr = lane_base + rt.slot * (uint32_t)T_STAGE_COLS;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 x = v[half * 4 + j];
                        const float hx = tf32_rna(x.x), hy = tf32_rna(x.y), hz = tf32_rna(x.z), hw = tf32_rna(x.w);
                        h[4 * j + 0] = __float_as_uint(hx);
                        h[4 * j + 1] = __float_as_uint(hy);
                        h[4 * j + 2] = __float_as_uint(hz);
                        h[4 * j + 3] = __float_as_uint(hw);
                        l[4 * j + 0] = __float_as_uint(tf32_rna(x.x - hx));
                        l[4 * j + 1] = __float_as_uint(tf32_rna(x.y - hy));
                        l[4 * j + 2] = __float_as_uint(tf32_rna(x.z - hz));
                        l[4 * j + 3] = __float_as_uint(tf32_rna(x.w - hw));
                    }
                    tmem_st16(taddr + half * 16, h);
                    tmem_st16(taddr + 32 + half * 16, l);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar0 + 8u * (BAR_A_EMPTY + ra.slot));  // the raw tile has been consumed
                    mbar_arrive(bar0 + 8u * (BAR_T_FULL + rt.slot));
                }
                if (q == 0 && lane == 0) dbg_mark(p, 3, dq);
            }
            ra.advance(p.nA);
            rt.advance(p.nT);
        }
        if (++nt == p.n_tiles) { nt = 0; ++mt; }
    }
}

template <bool DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_ts_kernel(const __grid_constant__ CUtensorMap tmA, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [raw A ring: nA x 16K][weight slots: nB x (hi | lo)][epilogue: 2 x (transpose + bias)][barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_slot_bytes = 2u * p.BN * 128u;
    const uint32_t a_off = 0;
    const uint32_t b_off = a_off + p.nA * (uint32_t)A_TILE_BYTES;
    const uint32_t epi_off = b_off + p.nB * b_slot_bytes;
    const uint32_t bar_off = epi_off + 2u * (uint32_t)p.epi_group_bytes;
    const uint32_t bar0 = base + bar_off;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8u * BAR_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nA; ++s) {
            mbar_init(bar0 + 8u * (BAR_A_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_A_EMPTY + s), 4);
        }
        for (int s = 0; s < p.nT; ++s) {
            mbar_init(bar0 + 8u * (BAR_T_FULL + s), 4);
            mbar_init(bar0 + 8u * (BAR_T_EMPTY + s), 1);
        }
        for (int s = 0; s < p.nB; ++s) {
            mbar_init(bar0 + 8u * (BAR_B_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_B_EMPTY + s), 1);
        }
        for (int s = 0; s < p.nAcc; ++s) {
            mbar_init(bar0 + 8u * (BAR_ACC_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_ACC_EMPTY + s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int item0 = blockIdx.x * p.items_per_cta;
    const int item1 = min(item0 + p.items_per_cta, p.total_items);

    // Everything above touched only this CTA's shared/tensor memory.  The weight producer may start right away (the
    // weight image is constant); every other role waits for the previous kernel of the stream (programmatic dependent
    // launch) before it reads activations / gates / residuals or writes the output.
    if (warp != NUM_THREADS / 32 - 1) pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===== A producer (the weight slots have their own warp, the last one: a full weight ring never holds
        // back the activation prefetch and vice versa) =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            Ring ra;
            uint32_t dq = 0;
            int mt = item0 / p.n_tiles, nt = item0 - mt * p.n_tiles;
            for (int item = item0; item < item1; ++item) {
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(bar0 + 8u * (BAR_A_EMPTY + ra.slot), ra.phase ^ 1u, 1);
                    mbar_expect_tx(bar0 + 8u * (BAR_A_FULL + ra.slot), A_TILE_BYTES);
                    tma_load_2d(base + a_off + ra.slot * (uint32_t)A_TILE_BYTES, &tmA, bar0 + 8u * (BAR_A_FULL + ra.slot),
                                kb * BK, mt * BM);
                    dbg_mark(p, 0, dq++);
                    ra.advance(p.nA);
                }
                if (++nt == p.n_tiles) { nt = 0; ++mt; }
            }
        }
    } else if (warp == NUM_THREADS / 32 - 1) {
        // ===== weight producer =====
        if (lane == 0) {
            Ring rb;
            int nt = item0 % p.n_tiles;
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wp);
            if (p.resident) {  // the whole weight image once, first-needed slots first
                const int total = p.n_tiles * p.KB;
                for (int i = 0; i < total; ++i) {
                    int s = nt * p.KB + i;
                    if (s >= total) s -= total;
                    mbar_expect_tx(bar0 + 8u * (BAR_B_FULL + s), b_slot_bytes);
                    bulk_load(base + b_off + s * b_slot_bytes, wsrc + (size_t)s * b_slot_bytes, b_slot_bytes,
                              bar0 + 8u * (BAR_B_FULL + s));
                }
            } else {
                for (int item = item0; item < item1; ++item) {
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(bar0 + 8u * (BAR_B_EMPTY + rb.slot), rb.phase ^ 1u, 0);
                        mbar_expect_tx(bar0 + 8u * (BAR_B_FULL + rb.slot), b_slot_bytes);
                        bulk_load(base + b_off + rb.slot * b_slot_bytes, wsrc + (size_t)(nt * p.KB + kb) * b_slot_bytes,
                                  b_slot_bytes, bar0 + 8u * (BAR_B_FULL + rb.slot));
                        rb.advance(p.nB);
                    }
                    if (++nt == p.n_tiles) nt = 0;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues) =====
        Ring rt, rb, racc;
        uint32_t dq = 0, b_seen = 0;
        const uint64_t desc0 = make_desc_sw128(base);
        const int k_tail = (p.K - (p.KB - 1) * BK + 7) / 8;
        int nt = item0 % p.n_tiles;
        for (int item = item0; item < item1; ++item) {
            mbar_wait(bar0 + 8u * (BAR_ACC_EMPTY + racc.slot), racc.phase ^ 1u, 2);
            const uint32_t d_tmem = tmem_base + racc.slot * (uint32_t)p.acc_stride;
            for (int kb = 0; kb < p.KB; ++kb, ++dq) {
                uint32_t slot;
                if (p.resident) {
                    slot = (uint32_t)(nt * p.KB + kb);
                    if (!((b_seen >> slot) & 1u)) {  // a resident slot lands once: no barrier round trip after that
                        mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), 0, 3);
                        b_seen |= 1u << slot;
                    }
                } else {
                    slot = rb.slot;
                    mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), rb.phase, 3);
                    rb.advance(p.nB);
                }
                const uint64_t dbh = desc0 + ((b_off + slot * b_slot_bytes) >> 4);
                const uint64_t dbl = dbh + ((p.BN * 128u) >> 4);
                const int ksteps = (kb == p.KB - 1) ? k_tail : BK / 8;
                mbar_wait(bar0 + 8u * (BAR_T_FULL + rt.slot), rt.phase, 4);
                tc_fence_after();
                if (lane == 0) dbg_mark(p, 4, dq);
                const uint32_t a_hi = tmem_base + (uint32_t)p.a_col0 + rt.slot * (uint32_t)T_STAGE_COLS;
                const uint32_t a_lo = a_hi + 32u;
                if (elect_one()) {
                    for (int k8 = 0; k8 < ksteps; ++k8) {
                        const uint64_t ko = (uint64_t)(k8 * 2);  // 32 bytes of K per step in the weight tile
                        const uint32_t ka = (uint32_t)(k8 * 8);  // 8 TMEM columns of K per step
                        umma_tf32_ts(d_tmem, a_lo + ka, dbh + ko, p.idesc, (kb | k8) ? 1u : 0u);  // small terms first
                        umma_tf32_ts(d_tmem, a_hi + ka, dbl + ko, p.idesc, 1u);
                        umma_tf32_ts(d_tmem, a_hi + ka, dbh + ko, p.idesc, 1u);
                    }
                    umma_commit(bar0 + 8u * (BAR_T_EMPTY + rt.slot));
                    if (!p.resident) umma_commit(bar0 + 8u * (BAR_B_EMPTY + slot));
                    if (kb == p.KB - 1) umma_commit(bar0 + 8u * (BAR_ACC_FULL + racc.slot));
                }
                __syncwarp();
                if (lane == 0) dbg_mark(p, 5, dq);
                rt.advance(p.nT);
            }
            racc.advance(p.nAcc);
            if (++nt == p.n_tiles) nt = 0;
        }
    } else if (warp < 2 + NUM_CONVERTERS / 32) {
        const int cw = warp - 2;
        if (p.scale != nullptr) converter_loop<true, DBG>(p, gbase + a_off, tmem_base, bar0, item0, item1, warp & 3, lane, cw >> 2);
        else converter_loop<false, DBG>(p, gbase + a_off, tmem_base, bar0, item0, item1, warp & 3, lane, cw >> 2);
    } else if (warp < 2 + (NUM_CONVERTERS + NUM_EPILOGUE) / 32) {
        // ===== epilogue =====
        const int ew8 = warp - (2 + NUM_CONVERTERS / 32);
        const int ew = ew8 & 3, eg = ew8 >> 2;
        float* stg = reinterpret_cast<float*>(gbase + epi_off + eg * p.epi_group_bytes) + ew * 32 * EPI_LD;
        float* s_bias = reinterpret_cast<float*>(gbase + epi_off + eg * p.epi_group_bytes + EPI_STAGE_BYTES);
        // every n tile's bias, once (named barrier 1+eg = the 4 warps of this epilogue group)
        for (int i = ew * 32 + lane; i < p.n_tiles * p.BN; i += 128) s_bias[i] = i < p.N ? __ldg(p.bias + i) : 0.0f;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        const bool has_res = p.res != nullptr;
        switch (p.act) {
            case YR_ACT_RELU6:
                if (has_res) epilogue_loop<YR_ACT_RELU6, true, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else if (p.up2) epilogue_loop<YR_ACT_RELU6, false, true, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else epilogue_loop<YR_ACT_RELU6, false, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                break;
            case YR_ACT_SWISH:
                if (has_res) epilogue_loop<YR_ACT_SWISH, true, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else if (p.up2) epilogue_loop<YR_ACT_SWISH, false, true, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else epilogue_loop<YR_ACT_SWISH, false, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                break;
            default:
                if (has_res) epilogue_loop<YR_ACT_NONE, true, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else if (p.up2) epilogue_loop<YR_ACT_NONE, false, true, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
                else epilogue_loop<YR_ACT_NONE, false, false, DBG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// W [K][N] row-major -> per (n tile, k block): [hi tile | lo tile], each BN rows (n) x 32 k-floats in the
// K-major SWIZZLE_128B image the MMA's B descriptor reads, zero padded (same format as pwconv_tc.cu).
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

struct Tiling {
    int BN, n_tiles, KB, nA, nT, nB, nAcc, resident, acc_stride, a_col0, epi_group_bytes;
    size_t smem;
};

static bool make_tiling(int K, int N, Tiling& t) {
    if (K <= 0 || N <= 0 || K % 8 || N % 4) return false;
    t.n_tiles = (N + 191) / 192;
    t.BN = ((N + t.n_tiles - 1) / t.n_tiles + 15) / 16 * 16;
    if (t.BN < 16) t.BN = 16;
    t.epi_group_bytes = EPI_STAGE_BYTES + (t.n_tiles * t.BN * 4 + 1023) / 1024 * 1024;  // + every n tile's bias
    t.KB = (K + BK - 1) / BK;
    t.acc_stride = (t.BN + 31) / 32 * 32;
    t.nAcc = t.acc_stride <= 64 ? 4 : 2;
    t.a_col0 = t.nAcc * t.acc_stride;
    t.nT = (512 - t.a_col0) / T_STAGE_COLS;
    if (t.nT > MAX_T) t.nT = MAX_T;
    if (t.nT < 2) return false;
    const long long slot = 2ll * t.BN * 128;
    const long long tile = A_TILE_BYTES;
    const long long fixed = 1024 + 2ll * t.epi_group_bytes + BAR_BYTES;
    const long long avail = SMEM_LIMIT - fixed;
    const long long wbytes = (long long)t.n_tiles * t.KB * slot;
    if (t.n_tiles * t.KB <= MAX_B && wbytes + 4 * tile <= avail) {
        t.resident = 1;
        t.nB = t.n_tiles * t.KB;
    } else {
        // streamed: the MMA issuer waits on weights (L2 latency ~2k cycles against ~800 cycles of MMA per slot), so the
        // weight ring gets the depth (up to 5 slots) and the raw A ring keeps 4 tiles
        t.resident = 0;
        long long nb = (avail - 4 * tile) / slot;  // measured: a 2-tile A ring with one more weight slot is slower
        if (nb > 5) nb = 5;
        if (nb > (long long)t.n_tiles * t.KB) nb = (long long)t.n_tiles * t.KB;
        if (nb < 2) nb = 2;
        t.nB = (int)nb;
    }
    long long na = (avail - t.nB * slot) / tile;
    if (na > MAX_A) na = MAX_A;
    // EVEN ring: the two converter groups alternate k-blocks, so with an even ring every slot belongs to one group
    // and that group sees each of the slot's mbarrier phases.  With an odd ring a group would skip the phase the
    // other group consumes, and a 1-bit parity wait cannot tell "two fills ago" from "this fill" when TMA loads
    // land out of order (observed on B200 with a 3-slot ring: stale tile read, then a protocol deadlock).
    na &= ~1ll;
    if (na < 2) return false;
    t.nA = (int)na;
    t.smem = (size_t)(fixed + t.nA * tile + t.nB * slot);
    return t.smem <= (size_t)SMEM_LIMIT;
}

}  // namespace ts

int launch_pw_ts(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w_tc && op.bias, "pw_ts: null pointer (w_tc = yr_pw_ts_pack output)");
    YR_CHECK_ARG(op.C > 0 && op.C % 8 == 0 && op.N > 0 && op.N % 8 == 0, "pw_ts: K=%d N=%d must be multiples of 8", op.C,
                 op.N);
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= op.N && op.ld_out % 4 == 0,
                 "pw_ts: bad ld_in=%d ld_out=%d", op.ld_in, op.ld_out);
    YR_CHECK_ARG(!op.res || (op.ld_res >= op.N && op.ld_res % 4 == 0), "pw_ts: bad ld_res=%d", op.ld_res);
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w_tc | (uintptr_t)op.bias | (uintptr_t)op.res |
                  (uintptr_t)op.scale) % 16 == 0, "pw_ts: pointers must be 16-byte aligned");
    const long long M = (long long)op.B * op.H * op.W;
    YR_CHECK_ARG(M > 0 && M < (1ll << 31) - 256, "pw_ts: bad row count");
    YR_CHECK_ARG((op.Ho == op.H && op.Wo == op.W) || (op.Ho == 2 * op.H && op.Wo == 2 * op.W && !op.res),
                 "pw_ts: output must be HxW, or 2Hx2W (fused nearest upsampling, no residual): got %dx%d for %dx%d", op.Ho, op.Wo,
                 op.H, op.W);
    ts::Tiling t;
    if (!ts::make_tiling(op.C, op.N, t)) {
        set_error("pw_ts: no tiling for K=%d N=%d", op.C, op.N);
        return YR_ERR_UNSUPPORTED;
    }
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) {
        set_error("pw_ts: cuTensorMapEncodeTiled is unavailable in this driver");
        return YR_ERR_CUDA;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)op.C, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)op.ld_in * 4};
    const cuuint32_t box[2] = {(cuuint32_t)ts::BK, (cuuint32_t)ts::BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(op.in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tc::l2_promotion(),
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("pw_ts: cuTensorMapEncodeTiled failed (%d) for K=%d M=%lld ld=%d", (int)cr, op.C, M, op.ld_in);
        return YR_ERR_CUDA;
    }
    ts::Params p;
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
    p.m_tiles = (int)((M + ts::BM - 1) / ts::BM);
    p.KB = t.KB;
    p.ld_out = op.ld_out;
    p.ld_res = op.ld_res;
    p.rows_per_img = op.H * op.W;
    p.img_w = op.W;
    p.up2 = (op.Ho == 2 * op.H && op.Wo == 2 * op.W) ? 1 : 0;
    p.act = op.act;
    p.nA = t.nA;
    p.nT = t.nT;
    p.nB = t.nB;
    p.nAcc = t.nAcc;
    p.resident = t.resident;
    p.acc_stride = t.acc_stride;
    p.a_col0 = t.a_col0;
    p.epi_group_bytes = t.epi_group_bytes;
    p.total_items = p.n_tiles * p.m_tiles;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(ts::pw_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) !=
                cudaSuccess ||
            cudaFuncSetAttribute(ts::pw_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) !=
                cudaSuccess) {
            set_error("pw_ts: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    const int max_ctas = tc::num_sms();
    // contiguous runs of m-major items per CTA: the n tiles of a 128-row block are mostly on one SM (the A re-read hits
    // L2 either way), and the split is item- not block-granular, which fills more SMs when there are few row blocks
    // (26x26 x 64 images = 338 blocks x 2 n tiles: 136 CTAs x 5 items instead of 113 x 6)
    p.items_per_cta = (p.total_items + max_ctas - 1) / max_ctas;
    const int grid = (p.total_items + p.items_per_cta - 1) / p.items_per_cta;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(t.BN >> 3) << 17) | ((uint32_t)(ts::BM >> 4) << 24);
    p.dbg = nullptr;
    static const bool debug = getenv("YR_PW_TC_DEBUG") != nullptr;  // developer aid only: timeline of CTA 0
    if (debug) {
        static long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc(&dbuf, 8 * ts::DBG_EV * sizeof(long long));
        cudaMemsetAsync(dbuf, 0, 8 * ts::DBG_EV * sizeof(long long), s);
        p.dbg = dbuf;
    }
    if (launch_pdl(debug ? ts::pw_ts_kernel<true> : ts::pw_ts_kernel<false>, dim3(grid), dim3(ts::NUM_THREADS), t.smem, s, tm,
                   p) != cudaSuccess) {
        set_error("pw_ts: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    if (debug) {
        static long long h[8 * ts::DBG_EV];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[8] = {"tma_issue", "a_full_seen", "t_empty_seen", "conv_done", "mma_start", "mma_issued",
                                "acc_full_seen", "epi_done"};
        long long t0 = h[0];
        fprintf(stderr, "pw_ts timeline K=%d N=%d BN=%d n_tiles=%d KB=%d nA=%d nT=%d nB=%d nAcc=%d resident=%d items/cta=%d grid=%d\n",
                p.K, p.N, p.BN, p.n_tiles, p.KB, p.nA, p.nT, p.nB, p.nAcc, p.resident, p.items_per_cta, grid);
        for (int r = 0; r < 8; ++r) {
            fprintf(stderr, "%-14s", names[r]);
            for (int i = 0; i < 40 && h[r * ts::DBG_EV + i]; ++i) fprintf(stderr, " %6lld", h[r * ts::DBG_EV + i] - t0);
            fprintf(stderr, "\n");
        }
    }
    return YR_OK;
}

}  // namespace yr

using namespace yr;

extern "C" int64_t yr_pw_ts_packed_floats(int K, int N) {
    ts::Tiling t;
    if (!ts::make_tiling(K, N, t)) return 0;
    return (int64_t)t.n_tiles * t.KB * 2 * t.BN * ts::BK;
}

extern "C" int yr_pw_ts_pack(const float* w, int K, int N, float* packed, void* stream) {
    YR_CHECK_ARG(w && packed, "pw_ts_pack: null pointer");
    ts::Tiling t;
    if (!ts::make_tiling(K, N, t)) {
        set_error("pw_ts_pack: no tensor-core tiling for K=%d N=%d", K, N);
        return YR_ERR_UNSUPPORTED;
    }
    YR_CHECK_ARG(((uintptr_t)packed) % 128 == 0, "pw_ts_pack: packed must be 128-byte aligned");
    const long long total = (long long)t.n_tiles * t.KB * t.BN * ts::BK;
    ts::pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, K, N, t.BN, t.n_tiles, t.KB, packed);
    YR_CHECK_LAUNCH("pw_ts_pack");
    return YR_OK;
}
