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
//
// DEPTHWISE FRONT (YR_OP_DWPW, template FRONT = stride 1 | 2): the same kernel with the 3x3 depthwise conv + BN +
// activation that PRECEDES the 1x1 conv computed by the converter warps, so the depthwise output - the widest tensor
// of an inverted-residual block (reference tf.keras.applications.MobileNetV2 blocks behind code/yolo3/override.py:339,
// MBConvBlock code/yolo3/efficientnet.py:501-522) - never exists in HBM: one write and one read of the 6x-expanded
// tensor less per block.  A work item is a TH x TW tile of output pixels of one image (<= 128 pixels = the 128 rows of
// the MMA); per 32-channel k-block the A producer brings the halo'd input box with ONE 4-D TMA (
// out-of-image coordinates zero-filled = TF 'SAME' padding); a converter group first computes the depthwise conv
// register-tiled like dw_tma_kernel (thread = 4 channels x 2 x 4 pixels, same operation order: bit-identical to running
// the two layers separately; taps and bias of the k-block ride at the end of the weight slot) into a 16 KB staging tile,
// then each thread takes one tile row (= its TMEM lane), splits it into (hi, lo) and writes tensor memory.
// The MMA issuer, weight producer and epilogue are unchanged apart from the row -> pixel mapping of the stores.
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
constexpr int BAR_BYTES = 2048;
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
    uint32_t a_slot_bytes, b_slot_bytes;  // ring slot sizes (depthwise front: halo'd box + this k-block's dw taps / weight slot)
    uint32_t a_taps_off;                  // depthwise front: byte offset of the taps + biases inside an activation slot
    // depthwise front: spatial tiling of the output (an item = one TH x TW tile of one image)
    int TH, TW, IW, tiles_h, tiles_w, Ho, Wo, pad_t, pad_l, dw_act;
    int conv_groups;  // converter groups taking k-blocks round-robin: 2, or 3 in the stride-1 depthwise front (the second
                      // epilogue group's warps convert instead: narrow outputs leave the epilogue idle, the depthwise
                      // phase is what bounds those layers); ring sizes are multiples of it
    int epi_groups;  // 2 or 1 (stride-2 depthwise front / three converter groups: the second group's resources go elsewhere)
    float* out2;     // stacked outputs (yr_op.aux): columns >= n_split go to out2 (row stride ld_out2), re-based to 0
    int ld_out2, n_split, first_linear;  // first_linear: columns < n_split skip the activation
    int epi_split;   // two groups: 1 = both drain EVERY tile, alternating its 32-column chunks (balanced for any tile count:
                     // the few tiles a CTA gets on the small layers rarely split evenly); 0 = the groups alternate tiles
    long long* dbg;
    int dbg_skip;  // debug instantiation only (YR_DWPW_SKIP): 1 = no depthwise math, 2 = no TMA box loads, 4 = no staging / split / TMEM stores, 8 = no MMAs
};
constexpr int DW_TAIL_BYTES = 2048;  // per slot of the packed image: 9 x 32 depthwise taps + 32 biases (1280 B), padded to 2 KB
constexpr int DW_TAPS_BYTES = 1280;  // what of it is copied: it travels WITH THE BOX into the activation slot, so the
                                     // converter groups never wait on the (shallow, MMA-paced) weight ring

constexpr int BAR_A_FULL = 0;
constexpr int BAR_A_EMPTY = BAR_A_FULL + MAX_A;
constexpr int BAR_T_FULL = BAR_A_EMPTY + MAX_A;
constexpr int BAR_T_EMPTY = BAR_T_FULL + MAX_T;
constexpr int BAR_B_FULL = BAR_T_EMPTY + MAX_T;
constexpr int BAR_B_EMPTY = BAR_B_FULL + MAX_B;
constexpr int BAR_ACC_FULL = BAR_B_EMPTY + MAX_B;
constexpr int BAR_ACC_EMPTY = BAR_ACC_FULL + MAX_ACC;
constexpr int BAR_BP_FULL = BAR_ACC_EMPTY + MAX_ACC;  // CTA pair: the PEER's half of a weight slot has landed (leader's copy)
constexpr int BAR_COUNT = BAR_BP_FULL + MAX_B;
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
// ---- CTA pair (cta_group::2) helpers: the two CTAs of a cluster run ONE tcgen05.mma over 256 rows; each holds half of
// every weight tile, so the per-CTA weight footprint halves (see the kernel comment) ----
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_addr), "r"(cta));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release, CTA scope) as CUTLASS' ClusterBarrier::arrive(cta_id) uses: the payload is tensor memory, ordered by
    // the tcgen05 fences; cluster-scope release/acquire would cost an L1 invalidate per k-block
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// waits of the leader on barriers that the peer CTA arrives on: the plain (CTA-scope) wait, as CUTLASS uses it for
// cta_group::2 pipelines - the cluster-scope acquire form made the pair kernel 35 % slower
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int what) { mbar_wait(bar, parity, what); }
__device__ __forceinline__ void umma_tf32_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {  // arrives on the barrier at this offset in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- epilogue (see pwconv_tc.cu for the access pattern; items here are m-major) ---------------
template <int ACT, bool HAS_RES, bool UP2, bool DBG, bool SP = false, int CG = 1>
__device__ __forceinline__ void epilogue_loop(const Params& p, float* stg, const float* s_bias, uint32_t tmem_base,
                                              uint32_t bar0, int item0, int item1, int q, int lane, int ewarp, int grp) {
    const uint32_t crank = CG == 2 ? cluster_rank() : 0u;  // CTA pair: item = a PAIR of row blocks, this CTA takes one
    const int sub_r = lane >> 3;
    const int sub_c = (lane & 7) << 2;
    uint32_t it = 0;
    Ring racc;
    int mt = item0 / p.n_tiles, nt = item0 - mt * p.n_tiles;
    for (int item = item0; item < item1; ++item, ++it) {
        const bool split = p.epi_groups == 2 && p.epi_split != 0;
        if (split || (p.epi_groups == 1 ? grp == 0 : (int)(it & 1u) == grp)) {
            const int cstep = split ? 64 : 32;
            const uint32_t acc = racc.slot;
            const int row0 = (CG == 2 ? mt * 2 + (int)crank : mt) * BM + q * 32;
            const int ncols = min(p.BN, p.N - nt * p.BN);
            const int rows = SP ? 32 : min(32, p.M - row0);
            // depthwise front: tile row -> output pixel of the [B, Ho, Wo] tensor (-1 = outside the tile / image)
            int pix[8];
            if (SP) {
                const int tw = mt % p.tiles_w, r2 = mt / p.tiles_w;
                const int th = r2 % p.tiles_h, bi = r2 / p.tiles_h;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = q * 32 + sub_r + 4 * i;
                    const int ty = r / p.TW, tx = r - ty * p.TW;
                    const int y = th * p.TH + ty, x = tw * p.TW + tx;
                    pix[i] = (ty < p.TH && y < p.Ho && x < p.Wo) ? (bi * p.Ho + y) * p.Wo + x : -1;
                }
            }
            const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.acc_stride;
            bool waited = false;
            float v[32];
            for (int c0 = split ? grp * 32 : 0; c0 < ncols; c0 += cstep) {
                const int c = c0 + sub_c;
                const int n = nt * p.BN + c;
                const bool col_ok = c < ncols;
                float4 rv[8];
                if (HAS_RES) {
                    const float* rp = p.res + (size_t)(row0 + sub_r) * p.ld_res + n;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (SP)
                            rv[i] = (col_ok && pix[i] >= 0) ? ldg4(p.res + (size_t)pix[i] * p.ld_res + n)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        else
                            rv[i] = (col_ok && sub_r + 4 * i < rows) ? ldg4(rp + (size_t)(4 * i) * p.ld_res)
                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
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
                if (!HAS_RES && c0 + cstep < ncols) tmem_ld32_issue(tsrc + c0 + cstep, v);
                __syncwarp();
                if (col_ok) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + n);
                    // stacked outputs: two 1x1 convs that read the same tensor run as ONE GEMM over [W1 | W2]; this
                    // lane's four columns belong to one of them (n_split % 4 == 0)
                    const bool second = !SP && !UP2 && p.n_split > 0 && n >= p.n_split;
                    const bool linear = !SP && !UP2 && p.n_split > 0 && !second && p.first_linear != 0;
                    const int ldo = second ? p.ld_out2 : p.ld_out;
                    float* op = (second ? p.out2 + (n - p.n_split) : p.out + n) + (size_t)(row0 + sub_r) * ldo;
                    const float* sp = stg + sub_r * EPI_LD + sub_c;
                    const size_t ostep = (size_t)4 * ldo;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (SP ? (pix[i] >= 0) : (sub_r + 4 * i < rows)) {
                            float4 x = *reinterpret_cast<const float4*>(sp + (4 * i) * EPI_LD);
                            if (HAS_RES || SP || UP2) {
                                // (kept in this exact form: the compiler contracts the activation's last multiply with
                                // the residual add into one FMA, and every instantiation must round the same way)
                                x.x = apply_act<ACT>(x.x + bv.x);
                                x.y = apply_act<ACT>(x.y + bv.y);
                                x.z = apply_act<ACT>(x.z + bv.z);
                                x.w = apply_act<ACT>(x.w + bv.w);
                            } else {  // (stacked outputs never carry a residual: checked by the launcher)
                                x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                                if (!linear) {
                                    x.x = apply_act<ACT>(x.x);
                                    x.y = apply_act<ACT>(x.y);
                                    x.z = apply_act<ACT>(x.z);
                                    x.w = apply_act<ACT>(x.w);
                                }
                            }
                            if (HAS_RES) { x.x += rv[i].x; x.y += rv[i].y; x.z += rv[i].z; x.w += rv[i].w; }
                            if (SP) {
                                st4(p.out + (size_t)pix[i] * p.ld_out + n, x);
                            } else if (!UP2) {
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
            if (CG == 2) mbar_arrive_cluster(map_to_cta(bar0 + 8u * (BAR_ACC_EMPTY + acc), 0));  // the leader issues the MMAs
            else mbar_arrive(bar0 + 8u * (BAR_ACC_EMPTY + acc));
            if (ewarp == 0 && lane == 0) dbg_mark(p, 7, it);
        }
        racc.advance(p.nAcc);
        if (++nt == p.n_tiles) { nt = 0; ++mt; }
    }
}

// ---- converters: raw fp32 A tile (smem) -> (hi, lo) TF32 columns of a TMEM stage ----------------
// Thread = tile row r = its TMEM lane.  The 128-byte row is read as 8 LDS.128 whose chunk index is
// XORed with (r & 7) (the TMA swizzle): the 8 lanes of a quarter-warp hit 8 different bank groups.
template <bool HAS_SCALE, bool DBG, int CG = 1>
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
                    const int grow = (CG == 2 ? mt * 2 + (int)cluster_rank() : mt) * BM + r;
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
                        l[4 * j + 0] = __float_as_uint(tf32_lo(x.x, hx));
                        l[4 * j + 1] = __float_as_uint(tf32_lo(x.y, hy));
                        l[4 * j + 2] = __float_as_uint(tf32_lo(x.z, hz));
                        l[4 * j + 3] = __float_as_uint(tf32_lo(x.w, hw));
                    }
                    tmem_st16(taddr + half * 16, h);
                    tmem_st16(taddr + 32 + half * 16, l);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar0 + 8u * (BAR_A_EMPTY + ra.slot));  // the raw tile has been consumed
                    if (CG == 2) mbar_arrive_cluster(map_to_cta(bar0 + 8u * (BAR_T_FULL + rt.slot), 0));
                    else mbar_arrive(bar0 + 8u * (BAR_T_FULL + rt.slot));
                }
                if (q == 0 && lane == 0) dbg_mark(p, 3, dq);
            }
            ra.advance(p.nA);
            rt.advance(p.nT);
        }
        if (++nt == p.n_tiles) { nt = 0; ++mt; }
    }
}

// ---- depthwise front: the converter group computes the 3x3 depthwise conv of the tile from the halo'd box ------
// Two phases per 32-channel k-block, both by the 128 threads of one converter group:
//   1. register-tiled depthwise conv exactly as dw_tma_kernel does it (thread = 4 channels x a 2 x 4 pixel patch:
//      24 (stride 2: 45) LDS.128 of the box for 32 outputs, 8 lanes cover a 128-byte pixel = conflict-free; same
//      operation order: acc = 0; acc = fma(x, w, acc) over (kh, kw) row-major; act(acc + bias)), results written to a
//      16 KB staging tile [128 pixels][32 channels] - the head of the box's own ring slot, once the whole group has read
//      the box - whose 16-byte chunks are XOR-swizzled with (row & 7);
//   2. the plain converter: thread = tile row = TMEM lane reads its 128-byte row (8 conflict-free LDS.128), splits
//      into (hi, lo) TF32 and writes tensor memory.
// The taps and bias of the k-block ride in the tail of the weight slot.  Named barrier 3 + group fences the staging tile.
template <int S, bool DBG>
__device__ __forceinline__ void dw_converter_loop(const Params& p, const uint8_t* a_ring, const uint8_t* b_ring,
                                                  uint32_t tmem_base, uint32_t bar0, int item0, int item1, int q, int lane,
                                                  int grp) {
    constexpr int PH = 2, PW = 4, KS = 3;
    constexpr int IN_ROWS = (PH - 1) * S + KS, IN_COLS = (PW - 1) * S + KS;
    Ring ra, rt;
    uint32_t dq = 0;
    const int gtid = q * 32 + lane;           // thread of the group; also the tile row / TMEM lane of phase 2
    const int cq = gtid & 7;                  // phase 1: channel quad (16-byte chunk of the pixel)
    const int pg = gtid >> 3;                 // phase 1: pixel patch
    const int gw = p.TW / PW;
    const int gx = pg % gw, gy = pg / gw;
    const bool worker = gy * PH < p.TH;       // tiles of fewer than 128 pixels leave the last patches / rows idle
    const bool live_row = gtid < p.TH * p.TW;
    const int box_off = worker ? ((gy * PH * S) * p.IW + gx * PW * S) * BK + cq * 4 : 0;  // floats
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0;
    const uint32_t bar_id = 3u + (uint32_t)grp;
    int turn = 0;  // dq % conv_groups
    for (int item = item0; item < item1; ++item) {
        for (int kb = 0; kb < p.KB; ++kb, ++dq) {
            const bool mine = turn == grp;
            if (++turn == p.conv_groups) turn = 0;
            if (mine) {
                mbar_wait(bar0 + 8u * (BAR_A_FULL + ra.slot), ra.phase, 5);
                if (q == 0 && lane == 0) dbg_mark(p, 1, dq);
                // ---- phase 1: depthwise conv of this thread's patch ----
                const float* sx = reinterpret_cast<const float*>(a_ring + (size_t)ra.slot * p.a_slot_bytes) + box_off;
                const float4* wd = reinterpret_cast<const float4*>(a_ring + (size_t)ra.slot * p.a_slot_bytes + p.a_taps_off) + cq;
                const float4 bv = wd[72];  // the accumulators start from the folded-BN bias (as in dw_tma_kernel)
                float4 acc[PH][PW];
#pragma unroll
                for (int t = 0; t < PH; ++t)
#pragma unroll
                    for (int o = 0; o < PW; ++o) acc[t][o] = bv;
                if (worker && !(DBG && (p.dbg_skip & 1))) {
#pragma unroll
                for (int rr = 0; rr < IN_ROWS; ++rr) {
                    float4 x[IN_COLS];
                    const float* row = sx + (size_t)rr * p.IW * BK;
#pragma unroll
                    for (int j = 0; j < IN_COLS; ++j) x[j] = *reinterpret_cast<const float4*>(row + j * BK);
#pragma unroll
                    for (int t = 0; t < PH; ++t) {
                        const int kh = rr - t * S;
                        if (kh < 0 || kh >= KS) continue;
#pragma unroll
                        for (int kw = 0; kw < KS; ++kw) {
                            const float4 ww = wd[(kh * KS + kw) * 8];
#pragma unroll
                            for (int o = 0; o < PW; ++o) {
                                fma4(acc[t][o], x[o * S + kw], ww);
                            }
                        }
                    }
                }
                }
                // the staging tile is the head of this box's OWN slot: every warp of the group must be done reading the
                // box before anyone overwrites it (no separate staging buffers: their 16 KB per group go to the A ring)
                float* stage = reinterpret_cast<float*>(const_cast<uint8_t*>(a_ring) + (size_t)ra.slot * p.a_slot_bytes);
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const bool skip2 = DBG && (p.dbg_skip & 4);
                if (worker && !skip2) {
#pragma unroll
                for (int t = 0; t < PH; ++t) {
#pragma unroll
                    for (int o = 0; o < PW; ++o) {
                        float4 v;
                        v.x = apply_act_rt(acc[t][o].x, p.dw_act);
                        v.y = apply_act_rt(acc[t][o].y, p.dw_act);
                        v.z = apply_act_rt(acc[t][o].z, p.dw_act);
                        v.w = apply_act_rt(acc[t][o].w, p.dw_act);
                        const int r = (gy * PH + t) * p.TW + gx * PW + o;
                        *reinterpret_cast<float4*>(stage + r * BK + ((cq ^ (r & 7)) << 2)) = v;
                    }
                }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                // ---- phase 2: row gtid of the staging tile -> (hi, lo) TF32 columns of the TMEM stage ----
                float4 v[8];
                if (!skip2) {
                    const float* rowp = stage + gtid * BK;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        v[c] = live_row ? *reinterpret_cast<const float4*>(rowp + ((c ^ (gtid & 7)) << 2))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                fence_proxy_async();  // generic-proxy writes to the slot are ordered before the TMA that refills it
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8u * (BAR_A_EMPTY + ra.slot));  // this warp is done with the slot
                mbar_wait(bar0 + 8u * (BAR_T_EMPTY + rt.slot), rt.phase ^ 1u, 7);
                tc_fence_after();
                if (q == 0 && lane == 0) dbg_mark(p, 2, dq);
                const uint32_t taddr = lane_base + rt.slot * (uint32_t)T_STAGE_COLS;
                if (!skip2) {
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
                        l[4 * j + 0] = __float_as_uint(tf32_lo(x.x, hx));
                        l[4 * j + 1] = __float_as_uint(tf32_lo(x.y, hy));
                        l[4 * j + 2] = __float_as_uint(tf32_lo(x.z, hz));
                        l[4 * j + 3] = __float_as_uint(tf32_lo(x.w, hw));
                    }
                    tmem_st16(taddr + half * 16, h);
                    tmem_st16(taddr + 32 + half * 16, l);
                }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar0 + 8u * (BAR_T_FULL + rt.slot));
                if (q == 0 && lane == 0) dbg_mark(p, 3, dq);
            }
            ra.advance(p.nA);
            rt.advance(p.nT);
        }
    }
}

// CG = 2: the CTA-PAIR form (plain front only).  The two CTAs of a cluster process a PAIR of 128-row blocks with ONE
// tcgen05.mma.cta_group::2 (M = 256) issued by the leader (cluster rank 0): each CTA converts its own rows into its own
// tensor memory and drains its own accumulator rows, but holds only HALF of every weight tile (the rows of its half of
// the n tile) - the per-CTA weight footprint halves, so layers whose (hi, lo) image does not fit one CTA's shared
// memory become resident (K = 256, N = 128: 256 KB -> 128 KB per CTA) and streamed layers get twice the ring depth.
// Cross-CTA signalling: the peer's converters and epilogue arrive on the LEADER's T_FULL / ACC_EMPTY barriers through
// the cluster address space, the peer's otherwise idle MMA warp relays "my half of weight slot s has landed" to the
// leader's BP_FULL barrier, and tcgen05.commit.cta_group::2 with a multicast mask releases T / B slots and publishes
// the accumulator in both CTAs at once.
template <bool DBG, int FRONT, int CG = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_ts_kernel(const __grid_constant__ CUtensorMap tmA, const Params p) {
    static_assert(CG == 1 || FRONT == 0, "the CTA-pair form exists for the plain front only");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [raw A ring: nA x 16K][weight slots: nB x (hi | lo)][epilogue: 2 x (transpose + bias)][barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_slot_bytes = p.b_slot_bytes;
    const uint32_t a_off = 0;
    const uint32_t b_off = a_off + p.nA * p.a_slot_bytes;
    const uint32_t epi_off = b_off + p.nB * b_slot_bytes;
    const uint32_t stage_off = epi_off + (uint32_t)(p.epi_groups * p.epi_group_bytes);
    const uint32_t bar_off = stage_off;
    const uint32_t bar0 = base + bar_off;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8u * BAR_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nA; ++s) {
            mbar_init(bar0 + 8u * (BAR_A_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_A_EMPTY + s), 4);
        }
        for (int s = 0; s < p.nT; ++s) {
            mbar_init(bar0 + 8u * (BAR_T_FULL + s), 4 * CG);   // CTA pair: the leader's copy collects both CTAs' converters
            mbar_init(bar0 + 8u * (BAR_T_EMPTY + s), 1);
        }
        for (int s = 0; s < p.nB; ++s) {
            mbar_init(bar0 + 8u * (BAR_B_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_B_EMPTY + s), 1);
            if (CG == 2) mbar_init(bar0 + 8u * (BAR_BP_FULL + s), 1);
        }
        for (int s = 0; s < p.nAcc; ++s) {
            mbar_init(bar0 + 8u * (BAR_ACC_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_ACC_EMPTY + s), 128 * CG * ((p.epi_groups == 2 && p.epi_split) ? 2 : 1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // both CTAs' barriers are initialised and both tensor memories allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t crank = CG == 2 ? cluster_rank() : 0u;

    // CTA pair: work is split over CLUSTERS (both CTAs walk the same items: pairs of row blocks x n tiles)
    const int item0 = (CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) * p.items_per_cta;
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
                int bx = 0, by = 0, bimg = 0;  // depthwise front: box origin of this item's tile (input coordinates)
                uint32_t box_bytes = A_TILE_BYTES;
                if constexpr (FRONT != 0) {
                    const int tw = mt % p.tiles_w, r2 = mt / p.tiles_w;
                    bx = tw * p.TW * FRONT - p.pad_l;
                    by = (r2 % p.tiles_h) * p.TH * FRONT - p.pad_t;
                    bimg = r2 / p.tiles_h;
                    box_bytes = (uint32_t)(((p.TH - 1) * FRONT + 3) * p.IW) * 128u;
                }
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(bar0 + 8u * (BAR_A_EMPTY + ra.slot), ra.phase ^ 1u, 1);
                    if (DBG && FRONT != 0 && (p.dbg_skip & 2)) {
                        mbar_expect_tx(bar0 + 8u * (BAR_A_FULL + ra.slot), 0);
                        dbg_mark(p, 0, dq++);
                        ra.advance(p.nA);
                        continue;
                    }
                    mbar_expect_tx(bar0 + 8u * (BAR_A_FULL + ra.slot), box_bytes + (FRONT != 0 ? (uint32_t)DW_TAPS_BYTES : 0u));
                    if constexpr (FRONT != 0) {
                        tma_load_4d(base + a_off + ra.slot * p.a_slot_bytes, &tmA, bar0 + 8u * (BAR_A_FULL + ra.slot), kb * BK,
                                    bx, by, bimg);
                        // this k-block's depthwise taps + biases, from behind its weight tiles in the packed image
                        bulk_load(base + a_off + ra.slot * p.a_slot_bytes + p.a_taps_off,
                                  reinterpret_cast<const uint8_t*>(p.wp) + (size_t)kb * (2u * p.BN * 128u + DW_TAIL_BYTES) +
                                      2u * p.BN * 128u,
                                  DW_TAPS_BYTES, bar0 + 8u * (BAR_A_FULL + ra.slot));
                    } else
                        tma_load_2d(base + a_off + ra.slot * p.a_slot_bytes, &tmA, bar0 + 8u * (BAR_A_FULL + ra.slot),
                                    kb * BK, (CG == 2 ? mt * 2 + (int)crank : mt) * BM);
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
            // one weight slot of the packed image = [hi tile: BN rows x 128 B | lo tile].  CTA pair: this CTA keeps rows
            // [rank * BN/2, +BN/2) of both tiles: local slot = [hi half | lo half] (two bulk copies, one barrier)
            const uint32_t g_slot = FRONT != 0 ? b_slot_bytes + DW_TAIL_BYTES   // (the taps behind the tiles go with the boxes)
                                               : (CG == 2 ? 2u * p.BN * 128u : b_slot_bytes);
            const uint32_t half = p.BN * 64u;
            auto load_slot = [&](uint32_t slot, int src_slot) {
                const uint32_t bar = bar0 + 8u * (BAR_B_FULL + slot);
                const uint32_t dst = base + b_off + slot * b_slot_bytes;
                const uint8_t* src = wsrc + (size_t)src_slot * g_slot;
                mbar_expect_tx(bar, b_slot_bytes);
                if (CG == 2) {
                    bulk_load(dst, src + crank * half, half, bar);
                    bulk_load(dst + half, src + 2u * half + crank * half, half, bar);
                } else {
                    bulk_load(dst, src, b_slot_bytes, bar);
                }
            };
            if (p.resident) {  // the whole weight image once, first-needed slots first
                const int total = p.n_tiles * p.KB;
                for (int i = 0; i < total; ++i) {
                    int s = nt * p.KB + i;
                    if (s >= total) s -= total;
                    load_slot((uint32_t)s, s);
                }
            } else {
                for (int item = item0; item < item1; ++item) {
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(bar0 + 8u * (BAR_B_EMPTY + rb.slot), rb.phase ^ 1u, 0);
                        load_slot(rb.slot, nt * p.KB + kb);
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
        if (CG == 2 && crank != 0) {
            // ----- peer CTA of a pair: no MMAs to issue; this warp relays "my half of weight slot s has landed" to the
            // leader, in the order the leader consumes the slots -----
            for (int item = item0; item < item1; ++item) {
                for (int kb = 0; kb < p.KB; ++kb) {
                    uint32_t slot;
                    bool relay = true;
                    if (p.resident) {
                        slot = (uint32_t)(nt * p.KB + kb);
                        relay = !((b_seen >> slot) & 1u);
                        if (relay) {
                            mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), 0, 3);
                            b_seen |= 1u << slot;
                        }
                    } else {
                        slot = rb.slot;
                        mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), rb.phase, 3);
                        rb.advance(p.nB);
                    }
                    if (relay && lane == 0) mbar_arrive_cluster(map_to_cta(bar0 + 8u * (BAR_BP_FULL + slot), 0));
                    __syncwarp();
                }
                if (++nt == p.n_tiles) nt = 0;
            }
        } else {
        for (int item = item0; item < item1; ++item) {
            if (CG == 2) mbar_wait_cluster(bar0 + 8u * (BAR_ACC_EMPTY + racc.slot), racc.phase ^ 1u, 2);
            else mbar_wait(bar0 + 8u * (BAR_ACC_EMPTY + racc.slot), racc.phase ^ 1u, 2);
            const uint32_t d_tmem = tmem_base + racc.slot * (uint32_t)p.acc_stride;
            for (int kb = 0; kb < p.KB; ++kb, ++dq) {
                uint32_t slot;
                if (p.resident) {
                    slot = (uint32_t)(nt * p.KB + kb);
                    if (!((b_seen >> slot) & 1u)) {  // a resident slot lands once: no barrier round trip after that
                        mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), 0, 3);
                        if (CG == 2) mbar_wait_cluster(bar0 + 8u * (BAR_BP_FULL + slot), 0, 3);
                        b_seen |= 1u << slot;
                    }
                } else {
                    slot = rb.slot;
                    mbar_wait(bar0 + 8u * (BAR_B_FULL + slot), rb.phase, 3);
                    if (CG == 2) mbar_wait_cluster(bar0 + 8u * (BAR_BP_FULL + slot), rb.phase, 3);
                    rb.advance(p.nB);
                }
                const uint64_t dbh = desc0 + ((b_off + slot * b_slot_bytes) >> 4);
                const uint64_t dbl = dbh + ((CG == 2 ? p.BN * 64u : p.BN * 128u) >> 4);  // CTA pair: half tiles per CTA
                const int ksteps = (kb == p.KB - 1) ? k_tail : BK / 8;
                if (CG == 2) mbar_wait_cluster(bar0 + 8u * (BAR_T_FULL + rt.slot), rt.phase, 4);
                else mbar_wait(bar0 + 8u * (BAR_T_FULL + rt.slot), rt.phase, 4);
                tc_fence_after();
                if (lane == 0) dbg_mark(p, 4, dq);
                const uint32_t a_hi = tmem_base + (uint32_t)p.a_col0 + rt.slot * (uint32_t)T_STAGE_COLS;
                const uint32_t a_lo = a_hi + 32u;
                if (elect_one()) {
                    for (int k8 = 0; k8 < ((DBG && (p.dbg_skip & 8)) ? 0 : ksteps); ++k8) {
                        const uint64_t ko = (uint64_t)(k8 * 2);  // 32 bytes of K per step in the weight tile
                        const uint32_t ka = (uint32_t)(k8 * 8);  // 8 TMEM columns of K per step
                        if (CG == 2) {
                            umma_tf32_ts_2cta(d_tmem, a_lo + ka, dbh + ko, p.idesc, (kb | k8) ? 1u : 0u);
                            umma_tf32_ts_2cta(d_tmem, a_hi + ka, dbl + ko, p.idesc, 1u);
                            umma_tf32_ts_2cta(d_tmem, a_hi + ka, dbh + ko, p.idesc, 1u);
                        } else {
                            umma_tf32_ts(d_tmem, a_lo + ka, dbh + ko, p.idesc, (kb | k8) ? 1u : 0u);  // small terms first
                            umma_tf32_ts(d_tmem, a_hi + ka, dbl + ko, p.idesc, 1u);
                            umma_tf32_ts(d_tmem, a_hi + ka, dbh + ko, p.idesc, 1u);
                        }
                    }
                    if (CG == 2) {
                        umma_commit_2cta(bar0 + 8u * (BAR_T_EMPTY + rt.slot));
                        if (!p.resident) umma_commit_2cta(bar0 + 8u * (BAR_B_EMPTY + slot));
                        if (kb == p.KB - 1) umma_commit_2cta(bar0 + 8u * (BAR_ACC_FULL + racc.slot));
                    } else {
                        umma_commit(bar0 + 8u * (BAR_T_EMPTY + rt.slot));
                        if (!p.resident) umma_commit(bar0 + 8u * (BAR_B_EMPTY + slot));
                        if (kb == p.KB - 1) umma_commit(bar0 + 8u * (BAR_ACC_FULL + racc.slot));
                    }
                }
                __syncwarp();
                if (lane == 0) dbg_mark(p, 5, dq);
                rt.advance(p.nT);
            }
            racc.advance(p.nAcc);
            if (++nt == p.n_tiles) nt = 0;
        }
        }
    } else if (warp < 2 + NUM_CONVERTERS / 32 ||
               (FRONT != 0 && p.conv_groups == 3 && warp >= 2 + NUM_CONVERTERS / 32 + 4 &&
                warp < 2 + (NUM_CONVERTERS + NUM_EPILOGUE) / 32)) {
        // converter warps; with three groups (depthwise front only, see Params::conv_groups) the second epilogue
        // group's warps are the third one - through the SAME call site, so the loop exists once in the instruction cache
        const int cw = warp < 2 + NUM_CONVERTERS / 32 ? warp - 2 : 8 + (warp & 3);
        if constexpr (FRONT != 0)
            dw_converter_loop<FRONT ? FRONT : 1, DBG>(p, gbase + a_off, gbase + b_off, tmem_base, bar0, item0, item1, warp & 3, lane,
                                                       cw >> 2);
        else if (p.scale != nullptr) converter_loop<true, DBG, CG>(p, gbase + a_off, tmem_base, bar0, item0, item1, warp & 3, lane, cw >> 2);
        else converter_loop<false, DBG, CG>(p, gbase + a_off, tmem_base, bar0, item0, item1, warp & 3, lane, cw >> 2);
    } else if (warp < 2 + (NUM_CONVERTERS + NUM_EPILOGUE) / 32) {
        // ===== epilogue =====
        const int ew8 = warp - (2 + NUM_CONVERTERS / 32);
        const int ew = ew8 & 3, eg = ew8 >> 2;
        float* stg = reinterpret_cast<float*>(gbase + epi_off + eg * p.epi_group_bytes) + ew * 32 * EPI_LD;
        float* s_bias = reinterpret_cast<float*>(gbase + epi_off + eg * p.epi_group_bytes + EPI_STAGE_BYTES);
        if (eg < p.epi_groups) {  // (a single-group launch leaves the second group's warps idle: it has no buffers)
        // every n tile's bias, once (named barrier 1+eg = the 4 warps of this epilogue group)
        for (int i = ew * 32 + lane; i < p.n_tiles * p.BN; i += 128) s_bias[i] = i < p.N ? __ldg(p.bias + i) : 0.0f;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        const bool has_res = p.res != nullptr;
#define YR_EPI(ACT_, RES_, UP_, SP_) \
    epilogue_loop<ACT_, RES_, UP_, DBG, SP_, CG>(p, stg, s_bias, tmem_base, bar0, item0, item1, warp & 3, lane, ew, eg)
        if constexpr (FRONT != 0) {  // depthwise front: rows are pixels of a spatial tile (no fused upsampling here)
            switch (p.act) {
                case YR_ACT_RELU6:
                    if (has_res) YR_EPI(YR_ACT_RELU6, true, false, true); else YR_EPI(YR_ACT_RELU6, false, false, true);
                    break;
                case YR_ACT_SWISH:
                    if (has_res) YR_EPI(YR_ACT_SWISH, true, false, true); else YR_EPI(YR_ACT_SWISH, false, false, true);
                    break;
                default:
                    if (has_res) YR_EPI(YR_ACT_NONE, true, false, true); else YR_EPI(YR_ACT_NONE, false, false, true);
            }
        } else {
            switch (p.act) {
                case YR_ACT_RELU6:
                    if (has_res) YR_EPI(YR_ACT_RELU6, true, false, false);
                    else if (p.up2) YR_EPI(YR_ACT_RELU6, false, true, false);
                    else YR_EPI(YR_ACT_RELU6, false, false, false);
                    break;
                case YR_ACT_SWISH:
                    if (has_res) YR_EPI(YR_ACT_SWISH, true, false, false);
                    else if (p.up2) YR_EPI(YR_ACT_SWISH, false, true, false);
                    else YR_EPI(YR_ACT_SWISH, false, false, false);
                    break;
                default:
                    if (has_res) YR_EPI(YR_ACT_NONE, true, false, false);
                    else if (p.up2) YR_EPI(YR_ACT_NONE, false, true, false);
                    else YR_EPI(YR_ACT_NONE, false, false, false);
            }
        }
#undef YR_EPI
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs may touch it
    if (warp == 1) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// W [K][N] row-major -> per (n tile, k block): [hi tile | lo tile], each BN rows (n) x 32 k-floats in the
// K-major SWIZZLE_128B image the MMA's B descriptor reads, zero padded (same format as pwconv_tc.cu).
__global__ void pack_kernel(const float* __restrict__ w, int K, int N, int BN, int n_tiles, int KB,
                            float* __restrict__ packed, int tail_floats) {
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
    const size_t slot_floats = (size_t)2 * BN * BK + tail_floats;
    const size_t off = (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 32 + (size_t)(((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);
    float* slot = packed + ((size_t)nt * KB + kb) * slot_floats;
    slot[off] = h;
    slot[(size_t)BN * BK + off] = l;
}

// depthwise front: per k-block the 9 x 32 taps ([tap][channel]) and the 32 biases behind the (hi | lo) weight tiles
__global__ void pack_dw_tail_kernel(const float* __restrict__ w_dw, const float* __restrict__ b_dw, int K, int KB,
                                    size_t slot_floats, size_t tail_off, float* __restrict__ packed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= KB * (DW_TAIL_BYTES / 4)) return;
    const int kb = idx / (DW_TAIL_BYTES / 4), e = idx % (DW_TAIL_BYTES / 4);
    float v = 0.f;
    const int k = kb * BK + (e & 31);
    if (k < K) {
        if (e < 288) v = w_dw[(size_t)(e >> 5) * K + k];
        else if (e < 320) v = b_dw[k];
    }
    packed[(size_t)kb * slot_floats + tail_off + e] = v;
}

inline int a_ring_for_streamed() {  // experiment knob YR_PW_STREAM_NA: raw A tiles kept when weights are streamed
    static const int v = [] {
        const char* e = getenv("YR_PW_STREAM_NA");
        const int r = e ? atoi(e) : 4;
        return r >= 2 && r <= 8 ? (r & ~1) : 4;
    }();
    return v;
}

struct Tiling {
    int BN, n_tiles, KB, nA, nT, nB, nAcc, resident, acc_stride, a_col0, epi_group_bytes;
    size_t smem;
    // depthwise front
    int TH, TW, IH, IW, tiles_h, tiles_w, epi_groups, conv_groups;
    uint32_t a_slot, b_slot;
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
    static const int res_na = [] {  // experiment knob: raw A tiles that must still fit beside a RESIDENT weight image
        const char* e = getenv("YR_PW_RESIDENT_NA");
        const int r = e ? atoi(e) : 4;
        return r >= 2 && r <= 8 ? (r & ~1) : 4;
    }();
    if (t.n_tiles * t.KB <= MAX_B && wbytes + res_na * tile <= avail) {
        t.resident = 1;
        t.nB = t.n_tiles * t.KB;
    } else {
        // streamed: the MMA issuer waits on weights (L2 latency ~2k cycles against ~800 cycles of MMA per slot), so the
        // weight ring gets the depth (up to 5 slots) and the raw A ring keeps 4 tiles
        t.resident = 0;
        long long nb = (avail - a_ring_for_streamed() * tile) / slot;  // measured: a 2-tile A ring with one more weight slot is slower
        if (nb > 6) nb = 6;
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

// CTA-pair form: same n tiles / k blocks / TMEM geometry (and the same packed weight image) as make_tiling, but a CTA
// holds half of every weight slot, so twice as much of the image fits (resident up to 2 x the single-CTA limit) and a
// streamed ring is twice as deep for the same bytes.
static bool make_tiling_pair(int K, int N, Tiling& t) {
    if (!make_tiling(K, N, t)) return false;
    if (t.BN % 16) return false;  // each CTA's half tile must be whole 8-row swizzle atoms
    const long long slot = (long long)t.BN * 128;  // [hi half | lo half]
    const long long tile = A_TILE_BYTES;
    static const int epi1_knob = [] {  // experiment knob: give one epilogue group's buffers to the weights when that makes
        const char* e = getenv("YR_PW_PAIR_EPI1");  // the image resident (long-K layers: the MMAs of a tile outlast its epilogue).
        return e ? atoi(e) : 1;                     // Measured: 52x52 K=256 -> N=128: 66.4 -> 62.3 us (0.65 of the HBM peak)
    }();
    t.epi_groups = 2;
    long long fixed = 1024 + 2ll * t.epi_group_bytes + BAR_BYTES;
    long long avail = SMEM_LIMIT - fixed;
    const long long wbytes = (long long)t.n_tiles * t.KB * slot;
    const bool fits2 = t.n_tiles * t.KB <= MAX_B && wbytes + 4 * tile <= avail;
    if (!fits2 && epi1_knob && t.KB >= 2 * ((t.BN + 31) / 32) && t.n_tiles * t.KB <= MAX_B &&
        wbytes + 4 * tile <= avail + t.epi_group_bytes) {
        t.epi_groups = 1;
        fixed -= t.epi_group_bytes;
        avail += t.epi_group_bytes;
    }
    if (t.n_tiles * t.KB <= MAX_B && wbytes + 4 * tile <= avail) {
        t.resident = 1;
        t.nB = t.n_tiles * t.KB;
    } else {
        t.resident = 0;
        long long nb = (avail - 4 * tile) / slot;
        if (nb > 8) nb = 8;
        if (nb > (long long)t.n_tiles * t.KB) nb = (long long)t.n_tiles * t.KB;
        if (nb < 2) return false;
        t.nB = (int)nb;
    }
    long long na = (avail - t.nB * slot) / tile;
    if (na > MAX_A) na = MAX_A;
    na &= ~1ll;
    if (na < 2) return false;
    t.nA = (int)na;
    t.b_slot = (uint32_t)slot;
    t.smem = (size_t)(fixed + t.nA * tile + t.nB * slot);
    return t.smem <= (size_t)SMEM_LIMIT;
}

constexpr long long DW_BOX_LIMIT = 76 * 1024;

// Depthwise front: TMEM / n-tile geometry as the plain kernel (single n tile only: the depthwise result is not
// recomputed per n tile), spatial tile = the TH x TW (<= 128 pixels) shape with the least halo + per-tile overhead.
static bool make_tiling_dw(int K, int N, int S, int Ho, int Wo, Tiling& t) {
    if (K <= 0 || N <= 0 || K % 8 || N % 4 || (S != 1 && S != 2) || Ho <= 0 || Wo <= 0) return false;
    t.n_tiles = (N + 191) / 192;
    if (t.n_tiles != 1) return false;
    t.BN = (N + 15) / 16 * 16;
    if (t.BN < 16) t.BN = 16;
    t.epi_group_bytes = EPI_STAGE_BYTES + (t.BN * 4 + 1023) / 1024 * 1024;
    t.KB = (K + BK - 1) / BK;
    t.acc_stride = (t.BN + 31) / 32 * 32;
    t.nAcc = t.acc_stride <= 64 ? 4 : 2;
    t.a_col0 = t.nAcc * t.acc_stride;
    t.nT = (512 - t.a_col0) / T_STAGE_COLS;
    if (t.nT > MAX_T) t.nT = MAX_T;
    if (t.nT < 2) return false;
    // tile = up to 128 pixels made of 2 x 4 pixel patches (one per converter thread and channel quad); the 64-pixel
    // shapes are the fallback when the boxes of a full tile leave no room for the weight slots (stride 2, wide layers)
    static const int shapes[9][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {4, 16}, {8, 8}, {16, 4}, {2, 32}};
    long long best = -1;
    Tiling bt = t;
    // A third converter group (the second epilogue group's warps) pays when the epilogue is light next to the depthwise
    // work of a tile - many k-blocks, few output columns - and costs when it is not (one epilogue group then drains every
    // tile).  Measured on B200 with the staging tile inside the A slot (ring of 6 boxes): 26x26x432->72 78 -> 68 us,
    // 13x13x720->120 39 -> 36 us, 104x104x144->24 211 -> 209 us, but 52x52x144->48 80 -> 82 us and 208x208x24->16 199 -> 260 us.
    static const int groups_knob = [] {  // experiment knob: 2 / 3 force the group count, 0 = the rule below
        const char* e = getenv("YR_DWPW_GROUPS");
        return e ? atoi(e) : 0;
    }();
    const int g_rule = t.KB >= 4 * ((t.BN + 31) / 32) ? 3 : 2;
    static const int dw_res_boxes = [] {  // experiment knob: 1 = resident weights even with one box per group
        const char* e = getenv("YR_DWPW_RES_BOXES");
        return e ? atoi(e) : 2;
    }();
    static const int dw_min_nb = [] {  // experiment knob: fewest weight slots accepted beside a double-depth box ring
        const char* e = getenv("YR_DWPW_MIN_NB");
        return e ? atoi(e) : 2;
    }();
    const int g_max = S == 1 ? (groups_knob == 2 || groups_knob == 3 ? groups_knob : g_rule) : 2;
    for (int i = 0; i < 9; ++i) {
        for (int G = g_max; G >= 2; --G) {
            Tiling c = t;
            const int TH = shapes[i][0], TW = shapes[i][1];
            c.TH = TH; c.TW = TW;
            c.IW = (TW - 1) * S + 3; c.IH = (TH - 1) * S + 3;
            if (c.IW > 256 || c.IH > 256 || (long long)c.IW * c.IH * 128 > DW_BOX_LIMIT) continue;
            c.a_slot = (uint32_t)(((long long)c.IW * c.IH * 128 + DW_TAPS_BYTES + 1023) / 1024 * 1024);  // box + taps
            c.b_slot = 2u * c.BN * 128u;
            c.conv_groups = G;
            // three converter groups borrow the second epilogue group's warps; stride 2: the boxes are ~4x the tile, so
            // one epilogue group gives its transpose buffers up
            c.epi_groups = (G == 3 || c.a_slot > 40 * 1024) ? 1 : 2;
            if (c.nT < G) continue;
            c.nT = c.nT / G * G;  // every ring is a multiple of the group count: a slot always belongs to one group,
                                  // which then sees every mbarrier phase of it (see make_tiling on even rings)
            const long long fixed = 1024 + (long long)c.epi_groups * c.epi_group_bytes + BAR_BYTES;  // (staging tiles live in the A slots)
            const long long avail = SMEM_LIMIT - fixed;
            // resident weights only if they leave room for two boxes per group (small boxes): a group whose next box is
            // requested only when it releases the current one waits the whole TMA latency (~3-5k cycles) every k-block
            const int res_boxes = (c.a_slot <= 32 * 1024 && dw_res_boxes == 2) ? 2 * G : G;
            const bool can_stream_deep = c.a_slot <= 32 * 1024 && (avail - 2ll * G * c.a_slot) / c.b_slot >= dw_min_nb;
            if (c.KB <= MAX_B && ((long long)c.KB * c.b_slot + (long long)res_boxes * c.a_slot <= avail ||
                                  (!can_stream_deep && (long long)c.KB * c.b_slot + (long long)G * c.a_slot <= avail))) {
                c.resident = 1;
                c.nB = c.KB;
            } else {
                c.resident = 0;
                // Only the MMA issuer consumes the weight ring (the taps travel with the boxes), so its depth is free of
                // the group count; two boxes per group (one being converted, one in flight) come first: a box takes
                // ~3-5k cycles to land, a weight slot is needed once per k-block period and refills in ~2.5k.
                long long nb = (avail - 2ll * G * c.a_slot) / c.b_slot;
                if (c.a_slot > 32 * 1024 || nb < dw_min_nb) nb = (avail - (long long)G * c.a_slot) / c.b_slot;
                if (nb > 6) nb = 6;
                if (nb < 2) continue;
                c.nB = (int)nb;
            }
            long long na = (avail - (long long)c.nB * c.b_slot) / c.a_slot;
            if (na > 6) na = 6;
            na = na / G * G;
            if (na < G) continue;
            c.nA = (int)na;
            c.smem = (size_t)(fixed + (long long)c.nA * c.a_slot + (long long)c.nB * c.b_slot);
            if (c.smem > (size_t)SMEM_LIMIT) continue;
            const long long tiles = (long long)((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
            const long long cost = tiles * ((long long)c.IW * c.IH + 96) * (G == 3 ? 2 : 3);  // 3 groups: ~1.5x the front rate
            if (best < 0 || cost < best) {
                best = cost;
                bt = c;
            }
        }
    }
    if (best < 0) return false;
    t = bt;
    t.tiles_h = (Ho + t.TH - 1) / t.TH;
    t.tiles_w = (Wo + t.TW - 1) / t.TW;
    return true;
}

}  // namespace ts

// Two epilogue groups can share the work two ways: alternate TILES (balanced when a CTA drains many tiles) or alternate
// the 32-column CHUNKS of every tile (balanced for any tile count, but 2:1 on a 96-column tile).  Pick the smaller
// makespan in columns; measured on B200: 26x26 48->288 (5 tiles per CTA) 24.5 -> 22.5 us, 13x13 120->720 (3 tiles)
// 24.5 -> 22.5 us with chunks; 208x208 24->96 fused (146 tiles of 3 chunks) 258 -> 309 us, so that one keeps tiles.
static int choose_epi_split(int items_per_cta, int ncols) {
    static const int knob = [] {  // experiment knob: YR_PW_EPI_SPLIT=0 / 1 force tiles / chunks
        const char* e = getenv("YR_PW_EPI_SPLIT");
        return e ? atoi(e) : -1;
    }();
    if (knob == 0 || knob == 1) return knob;
    long long g0 = 0, g1 = 0;
    for (int c = 0; c < ncols; c += 32) ((c >> 5) & 1 ? g1 : g0) += (ncols - c < 32 ? ncols - c : 32);
    const long long by_chunks = (long long)items_per_cta * (g0 > g1 ? g0 : g1);
    const long long by_tiles = (long long)((items_per_cta + 1) / 2) * ncols;
    return by_chunks < by_tiles ? 1 : 0;
}

// Fused depthwise 3x3 (+BN +act) -> pointwise 1x1 (+BN +act +residual): YR_OP_DWPW.
int launch_dwpw(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w_tc && op.bias, "dwpw: null pointer (w_tc = yr_dwpw_pack output)");
    YR_CHECK_ARG(op.k == 3 && (op.stride == 1 || op.stride == 2), "dwpw: depthwise must be 3x3, stride 1 or 2");
    YR_CHECK_ARG(op.C > 0 && op.C % 8 == 0 && op.N > 0 && op.N % 8 == 0, "dwpw: C=%d N=%d must be multiples of 8", op.C, op.N);
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= op.N && op.ld_out % 4 == 0,
                 "dwpw: bad ld_in=%d ld_out=%d", op.ld_in, op.ld_out);
    YR_CHECK_ARG(!op.res || (op.ld_res >= op.N && op.ld_res % 4 == 0), "dwpw: bad ld_res=%d", op.ld_res);
    YR_CHECK_ARG(!op.scale, "dwpw: an SE gate between the depthwise and the pointwise conv cannot be fused");
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w_tc | (uintptr_t)op.bias | (uintptr_t)op.res) % 16 == 0,
                 "dwpw: pointers must be 16-byte aligned");
    YR_CHECK_ARG(op.B > 0 && op.H > 0 && op.W > 0 && op.Ho > 0 && op.Wo > 0 &&
                     (long long)op.B * op.Ho * op.Wo * (op.ld_out > op.ld_res ? op.ld_out : op.ld_res) < (1ll << 31),
                 "dwpw: bad geometry");
    YR_CHECK_ARG(op.mode >= YR_ACT_NONE && op.mode <= YR_ACT_SWISH, "dwpw: mode = activation of the depthwise conv");
    ts::Tiling t;
    if (!ts::make_tiling_dw(op.C, op.N, op.stride, op.Ho, op.Wo, t)) {
        set_error("dwpw: no tiling for C=%d N=%d stride %d %dx%d", op.C, op.N, op.stride, op.Ho, op.Wo);
        return YR_ERR_UNSUPPORTED;
    }
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) {
        set_error("dwpw: cuTensorMapEncodeTiled is unavailable in this driver");
        return YR_ERR_CUDA;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[4] = {(cuuint64_t)op.C, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)op.ld_in * 4, (cuuint64_t)op.W * op.ld_in * 4,
                                (cuuint64_t)op.H * op.W * op.ld_in * 4};
    const cuuint32_t box[4] = {(cuuint32_t)ts::BK, (cuuint32_t)t.IW, (cuuint32_t)t.IH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(op.in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tc::l2_promotion(),
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("dwpw: cuTensorMapEncodeTiled failed (%d) for C=%d H=%d W=%d ld=%d box %dx%d", (int)cr, op.C, op.H, op.W,
                  op.ld_in, t.IW, t.IH);
        return YR_ERR_CUDA;
    }
    ts::Params p = {};
    p.wp = op.w_tc;
    p.bias = op.bias;
    p.res = op.res;
    p.scale = nullptr;
    p.out = (float*)op.out;
    p.M = op.B * t.tiles_h * t.tiles_w * ts::BM;
    p.K = op.C;
    p.N = op.N;
    p.BN = t.BN;
    p.n_tiles = 1;
    p.m_tiles = op.B * t.tiles_h * t.tiles_w;
    p.KB = t.KB;
    p.ld_out = op.ld_out;
    p.ld_res = op.ld_res;
    p.rows_per_img = op.Ho * op.Wo;
    p.img_w = op.Wo;
    p.up2 = 0;
    p.act = op.act;
    p.nA = t.nA;
    p.nT = t.nT;
    p.nB = t.nB;
    p.nAcc = t.nAcc;
    p.resident = t.resident;
    p.acc_stride = t.acc_stride;
    p.a_col0 = t.a_col0;
    p.epi_group_bytes = t.epi_group_bytes;
    p.total_items = p.m_tiles;
    p.a_slot_bytes = t.a_slot;
    p.b_slot_bytes = t.b_slot;
    p.a_taps_off = (uint32_t)((t.IW * t.IH * 128 + 127) / 128 * 128);
    p.out2 = nullptr; p.ld_out2 = 0; p.n_split = 0; p.first_linear = 0;
    p.TH = t.TH; p.TW = t.TW; p.IW = t.IW; p.tiles_h = t.tiles_h; p.tiles_w = t.tiles_w;
    p.Ho = op.Ho; p.Wo = op.Wo; p.pad_t = op.pad_t; p.pad_l = op.pad_l; p.dw_act = op.mode;
    p.epi_groups = t.epi_groups;
    p.epi_split = 0;  // set with items_per_cta below
    p.conv_groups = t.conv_groups;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(ts::pw_ts_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) != cudaSuccess ||
            cudaFuncSetAttribute(ts::pw_ts_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) != cudaSuccess ||
            cudaFuncSetAttribute(ts::pw_ts_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) != cudaSuccess ||
            cudaFuncSetAttribute(ts::pw_ts_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) != cudaSuccess) {
            set_error("dwpw: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    const int max_ctas = tc::num_sms();
    p.items_per_cta = (p.total_items + max_ctas - 1) / max_ctas;
    p.epi_split = p.epi_groups == 2 ? choose_epi_split(p.items_per_cta, p.BN < p.N ? p.BN : p.N) : 0;
    const int grid = (p.total_items + p.items_per_cta - 1) / p.items_per_cta;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(t.BN >> 3) << 17) | ((uint32_t)(ts::BM >> 4) << 24);
    p.dbg = nullptr;
    p.dbg_skip = 0;
    static const bool debug = getenv("YR_PW_TC_DEBUG") != nullptr;  // developer aid only: timeline of CTA 0
    // YR_PW_TC_DEBUG=2: the debug instantiation with its YR_DWPW_SKIP knobs but no timeline and no synchronisation,
    // so that event timings around the launch stay valid
    static const bool timeline = debug && atoi(getenv("YR_PW_TC_DEBUG")) != 2;
    if (debug) {
        if (timeline) {
            static long long* dbuf = nullptr;
            if (!dbuf) cudaMalloc(&dbuf, 8 * ts::DBG_EV * sizeof(long long));
            cudaMemsetAsync(dbuf, 0, 8 * ts::DBG_EV * sizeof(long long), s);
            p.dbg = dbuf;
        }
        const char* e = getenv("YR_DWPW_SKIP");
        p.dbg_skip = e ? atoi(e) : 0;
    }
    auto kern = op.stride == 1 ? (debug ? ts::pw_ts_kernel<true, 1> : ts::pw_ts_kernel<false, 1>)
                               : (debug ? ts::pw_ts_kernel<true, 2> : ts::pw_ts_kernel<false, 2>);
    if (launch_pdl(kern, dim3(grid), dim3(ts::NUM_THREADS), t.smem, s, tm, p) != cudaSuccess) {
        set_error("dwpw: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    if (timeline) {
        static long long h[8 * ts::DBG_EV];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[8] = {"tma_issue", "a_full_seen", "t_empty_seen", "conv_done", "mma_start", "mma_issued",
                                "acc_full_seen", "epi_done"};
        long long t0 = h[0];
        fprintf(stderr, "dwpw timeline groups=%d C=%d N=%d s%d %dx%d tile %dx%d box %dx%d tiles %dx%d BN=%d KB=%d nA=%d nT=%d nB=%d nAcc=%d resident=%d items/cta=%d grid=%d smem=%zu\n",
                p.conv_groups, p.K, p.N, op.stride, op.Ho, op.Wo, t.TH, t.TW, t.IH, t.IW, t.tiles_h, t.tiles_w, p.BN, p.KB, p.nA, p.nT, p.nB,
                p.nAcc, p.resident, p.items_per_cta, grid, t.smem);
        for (int r = 0; r < 8; ++r) {
            fprintf(stderr, "%-14s", names[r]);
            for (int i = 0; i < 40 && h[r * ts::DBG_EV + i]; ++i) fprintf(stderr, " %6lld", h[r * ts::DBG_EV + i] - t0);
            fprintf(stderr, "\n");
        }
    }
    return YR_OK;
}

template <int CG>
static int launch_pw_ts_cg(const yr_op& op, cudaStream_t s) {
    YR_CHECK_ARG(op.in && op.out && op.w_tc && op.bias, "pw_ts: null pointer (w_tc = yr_pw_ts_pack output)");
    YR_CHECK_ARG(op.C > 0 && op.C % 8 == 0 && op.N > 0 && op.N % 8 == 0, "pw_ts: K=%d N=%d must be multiples of 8", op.C,
                 op.N);
    const int n_first = op.K2 > 0 ? op.K2 : op.N;  // stacked outputs: columns [0, K2) -> out, [K2, N) -> aux
    YR_CHECK_ARG(op.ld_in >= op.C && op.ld_in % 4 == 0 && op.ld_out >= n_first && op.ld_out % 4 == 0,
                 "pw_ts: bad ld_in=%d ld_out=%d", op.ld_in, op.ld_out);
    YR_CHECK_ARG(!op.res || (op.ld_res >= op.N && op.ld_res % 4 == 0), "pw_ts: bad ld_res=%d", op.ld_res);
    YR_CHECK_ARG(((uintptr_t)op.in | (uintptr_t)op.out | (uintptr_t)op.w_tc | (uintptr_t)op.bias | (uintptr_t)op.res |
                  (uintptr_t)op.scale) % 16 == 0, "pw_ts: pointers must be 16-byte aligned");
    YR_CHECK_ARG(op.K2 == 0 || (op.K2 > 0 && op.K2 < op.N && op.K2 % 4 == 0 && op.aux && (uintptr_t)op.aux % 16 == 0 &&
                                op.ld_in2 >= op.N - op.K2 && op.ld_in2 % 4 == 0 && !op.res && op.Ho == op.H && op.Wo == op.W),
                 "pw_ts: stacked outputs need 0 < K2 < N, K2 %% 4 == 0, aux (second output, 16-byte aligned), ld_in2 >= N - K2, "
                 "no residual and no fused upsampling (K2=%d N=%d ld_in2=%d)", op.K2, op.N, op.ld_in2);
    const long long M = (long long)op.B * op.H * op.W;
    YR_CHECK_ARG(M > 0 && M < (1ll << 31) - 256, "pw_ts: bad row count");
    YR_CHECK_ARG((op.Ho == op.H && op.Wo == op.W) || (op.Ho == 2 * op.H && op.Wo == 2 * op.W && !op.res),
                 "pw_ts: output must be HxW, or 2Hx2W (fused nearest upsampling, no residual): got %dx%d for %dx%d", op.Ho, op.Wo,
                 op.H, op.W);
    ts::Tiling t;
    if (!(CG == 2 ? ts::make_tiling_pair(op.C, op.N, t) : ts::make_tiling(op.C, op.N, t))) {
        set_error("pw_ts: no tiling for K=%d N=%d (cta_group %d)", op.C, op.N, CG);
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
    p.out2 = op.K2 > 0 ? op.aux : nullptr;
    p.ld_out2 = op.ld_in2;
    p.n_split = op.K2 > 0 ? op.K2 : 0;
    p.first_linear = op.K3;
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
    p.total_items = p.n_tiles * (CG == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles);  // CTA pair: items are PAIRS of row blocks
    p.a_slot_bytes = ts::A_TILE_BYTES;
    p.b_slot_bytes = CG == 2 ? t.b_slot : 2u * t.BN * 128u;
    p.TH = p.TW = p.IW = p.tiles_h = p.tiles_w = 1;
    p.Ho = p.Wo = p.pad_t = p.pad_l = p.dw_act = 0;
    p.epi_groups = CG == 2 ? t.epi_groups : 2;
    p.epi_split = 0;  // set with items_per_cta below
    p.conv_groups = 2;
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(ts::pw_ts_kernel<false, 0, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) !=
                cudaSuccess ||
            cudaFuncSetAttribute(ts::pw_ts_kernel<true, 0, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::SMEM_LIMIT) !=
                cudaSuccess) {
            set_error("pw_ts: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    const int max_ctas = tc::num_sms() / CG;  // CTA pair: work units are clusters of two CTAs
    // contiguous runs of m-major items per CTA: the n tiles of a 128-row block are mostly on one SM (the A re-read hits
    // L2 either way), and the split is item- not block-granular, which fills more SMs when there are few row blocks
    // (26x26 x 64 images = 338 blocks x 2 n tiles: 136 CTAs x 5 items instead of 113 x 6)
    p.items_per_cta = (p.total_items + max_ctas - 1) / max_ctas;
    p.epi_split = p.epi_groups == 2 ? choose_epi_split(p.items_per_cta, p.BN < p.N ? p.BN : p.N) : 0;
    const int grid = CG * ((p.total_items + p.items_per_cta - 1) / p.items_per_cta);
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(t.BN >> 3) << 17) | ((uint32_t)((CG * ts::BM) >> 4) << 24);
    p.dbg = nullptr;
    p.dbg_skip = 0;
    static const bool debug = getenv("YR_PW_TC_DEBUG") != nullptr;  // developer aid only: timeline of CTA 0
    if (debug) {
        static long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc(&dbuf, 8 * ts::DBG_EV * sizeof(long long));
        cudaMemsetAsync(dbuf, 0, 8 * ts::DBG_EV * sizeof(long long), s);
        p.dbg = dbuf;
    }
    if (CG == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(ts::NUM_THREADS);
        cfg.dynamicSmemBytes = t.smem;
        cfg.stream = s;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (cudaLaunchKernelEx(&cfg, ts::pw_ts_kernel<false, 0, CG>, tm, p) != cudaSuccess) {
            set_error("pw_ts (CTA pair): launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
    } else if (launch_pdl(debug ? ts::pw_ts_kernel<true, 0, CG> : ts::pw_ts_kernel<false, 0, CG>, dim3(grid), dim3(ts::NUM_THREADS),
                          t.smem, s, tm, p) != cudaSuccess) {
        set_error("pw_ts: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    if (debug && CG == 1) {
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

int launch_pw_ts(const yr_op& op, cudaStream_t s) { return launch_pw_ts_cg<1>(op, s); }
int launch_pw_ts2(const yr_op& op, cudaStream_t s) { return launch_pw_ts_cg<2>(op, s); }

}  // namespace yr

using namespace yr;

/* 1 when the CTA-pair kernel (variant 4) has a tiling for a K x N layer (same weight image as variant 3). */
extern "C" int yr_pw_ts2_supported(int K, int N) {
    ts::Tiling t;
    return ts::make_tiling_pair(K, N, t) ? 1 : 0;
}

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
    ts::pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, K, N, t.BN, t.n_tiles, t.KB, packed, 0);
    YR_CHECK_LAUNCH("pw_ts_pack");
    return YR_OK;
}

// Weight image of a fused depthwise -> pointwise pair (YR_OP_DWPW): per 32-channel k-block the pointwise (hi | lo)
// TF32 tiles of yr_pw_ts_pack followed by that k-block's depthwise taps [9][32] and biases [32] (BN folded).
// The image depends on N and K only (not on the spatial size).
extern "C" int64_t yr_dwpw_packed_floats(int K, int N) {
    ts::Tiling t;
    if (!ts::make_tiling_dw(K, N, 1, 8, 16, t)) return 0;
    return (int64_t)t.KB * (2 * t.BN * ts::BK + ts::DW_TAIL_BYTES / 4);
}

extern "C" int yr_dwpw_pack(const float* w_pw, int K, int N, const float* w_dw, const float* b_dw, float* packed,
                            void* stream) {
    YR_CHECK_ARG(w_pw && w_dw && b_dw && packed, "dwpw_pack: null pointer");
    ts::Tiling t;
    if (!ts::make_tiling_dw(K, N, 1, 8, 16, t)) {
        set_error("dwpw_pack: no fused tiling for K=%d N=%d", K, N);
        return YR_ERR_UNSUPPORTED;
    }
    YR_CHECK_ARG(((uintptr_t)packed) % 128 == 0, "dwpw_pack: packed must be 128-byte aligned");
    const int tail = ts::DW_TAIL_BYTES / 4;
    const long long total = (long long)t.KB * t.BN * ts::BK;
    ts::pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_pw, K, N, t.BN, 1, t.KB, packed, tail);
    YR_CHECK_LAUNCH("dwpw_pack");
    const int n2 = t.KB * tail;
    ts::pack_dw_tail_kernel<<<(n2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        w_dw, b_dw, K, t.KB, (size_t)2 * t.BN * ts::BK + tail, (size_t)2 * t.BN * ts::BK, packed);
    YR_CHECK_LAUNCH("dwpw_pack");
    return YR_OK;
}

/* 1 when (C, N, stride, Ho, Wo) has a fused depthwise->pointwise tiling, else 0 (the engine then runs the two ops). */
extern "C" int yr_dwpw_supported(int C, int N, int stride, int Ho, int Wo) {
    ts::Tiling t;
    return ts::make_tiling_dw(C, N, stride, Ho, Wo, t) ? 1 : 0;
}

/* The plan the fused kernel would run for this geometry (host-only query, no GPU needed): plan[16] = {tile rows, tile
 * cols, box rows, box cols, tiles per image (rows), (cols), converter groups, epilogue groups, box ring, TMEM stage ring,
 * weight ring, accumulators, weights resident, n tile width, k-blocks, dynamic shared memory in bytes}.  1 / 0 as
 * yr_dwpw_supported. */
extern "C" int yr_dwpw_plan(int C, int N, int stride, int Ho, int Wo, int32_t* plan) {
    ts::Tiling t;
    if (!plan || !ts::make_tiling_dw(C, N, stride, Ho, Wo, t)) return 0;
    const int v[16] = {t.TH, t.TW, t.IH, t.IW, t.tiles_h, t.tiles_w, t.conv_groups, t.epi_groups, t.nA, t.nT, t.nB, t.nAcc,
                       t.resident, t.BN, t.KB, (int)t.smem};
    for (int i = 0; i < 16; ++i) plan[i] = v[i];
    return 1;
}
