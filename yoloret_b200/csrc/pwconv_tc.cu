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
#include "tc_common.cuh"

namespace yr {
namespace tc {

constexpr int BM = 128;                      // rows per tile = UMMA M
constexpr int BK = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 4;    // 16 KB (raw/hi) ; lo tile has the same size
constexpr int NUM_CONVERTERS = 256;          // converter threads: 2 groups of 4 warps, alternating k-blocks
constexpr int NUM_EPILOGUE = 256;            // epilogue threads: 2 groups of 4 warps, alternating tiles
constexpr int NUM_THREADS = 64 + NUM_CONVERTERS + NUM_EPILOGUE;  // + producer warp + MMA warp
constexpr int EPI_LD = 36;                   // padded row (floats) of the per-warp transpose buffer
constexpr int EPI_GROUP_BYTES = 4 * 32 * EPI_LD * 4 + 1024;  // per epilogue group: 4 transpose buffers + the n tile's bias
constexpr int SMEM_LIMIT = 232448;           // 227 KB per CTA
constexpr int MAX_A_STAGES = 12, MAX_L_SLOTS = 4, MAX_B_SLOTS = 24, MAX_ACC = 8;
constexpr int BAR_BYTES = 1024;

struct Params {
    const float* wp;      // packed weight image, see pack_kernel
    const float* bias;
    const float* res;
    const float* scale;
    float* out;
    int M, K, N, BN, n_tiles, m_tiles, KB;
    int ld_out, ld_res, rows_per_img, act;
    int up2, img_w;  // up2: every output row is stored to its 2x2 nearest-upsampled pixels (fused UpSampling2D)
    int nA, nL, nB, nAcc, nEpi, resident, tmem_cols, acc_stride, items_per_cta, total_items;
    uint32_t idesc;
    long long* dbg;  // optional timeline of CTA 0 (YR_PW_TC_DEBUG=1), else NULL
};

// ---- barrier table (8 bytes each, at the end of the dynamic shared memory) -------------------
constexpr int BAR_A_FULL = 0;                          // [MAX_A_STAGES]  TMA landed a raw A tile
constexpr int BAR_A_CONV = BAR_A_FULL + MAX_A_STAGES;  // [MAX_A_STAGES]  converters wrote (hi, lo)
constexpr int BAR_A_EMPTY = BAR_A_CONV + MAX_A_STAGES; // [MAX_A_STAGES]  MMAs that read the slot retired
constexpr int BAR_L_EMPTY = BAR_A_EMPTY + MAX_A_STAGES;  // [MAX_L_SLOTS]
constexpr int BAR_B_FULL = BAR_L_EMPTY + MAX_L_SLOTS;  // [MAX_B_SLOTS]
constexpr int BAR_B_EMPTY = BAR_B_FULL + MAX_B_SLOTS;  // [MAX_B_SLOTS]
constexpr int BAR_ACC_FULL = BAR_B_EMPTY + MAX_B_SLOTS;  // [MAX_ACC]
constexpr int BAR_ACC_EMPTY = BAR_ACC_FULL + MAX_ACC;  // [MAX_ACC]
constexpr int BAR_COUNT = BAR_ACC_EMPTY + MAX_ACC;
static_assert(BAR_COUNT * 8 + 8 <= BAR_BYTES, "barrier table overflows its reservation");

constexpr int DBG_EV = 256;  // events per role in the debug timeline
// compiled in only for the debug instantiation of the kernel (YR_PW_TC_DEBUG=1); see pwconv_ts.cu
template <bool DBG>
__device__ __forceinline__ void dbg_mark_t(const Params& p, int role, uint32_t idx) {
    if (DBG) {
        if (p.dbg != nullptr && blockIdx.x == 0 && idx < DBG_EV) p.dbg[role * DBG_EV + idx] = clock64();
    }
}
#define dbg_mark(p, role, idx) dbg_mark_t<DBG>(p, role, idx)

// Streamed-weight layers walk their tiles in GROUPS of two consecutive M tiles of one n tile: each weight slot
// is fetched once per group and feeds the k-block of both tiles, which halves the L2 -> SM weight traffic that
// bounds the wide head layers.  (Resident-weight layers: groups of one.)
__device__ __forceinline__ int group_size(const Params& p, int item, int item1, int mt) {
    // needs >= 4 accumulators: with 2 (N > 128) a pair would own all of TMEM and serialise MMA and epilogue
    return (!p.resident && p.nAcc >= 4 && item + 1 < item1 && mt + 1 < p.m_tiles) ? 2 : 1;
}

// ---- epilogue: TMEM -> registers -> smem transpose -> bias / activation / residual -> 128-bit row stores
// One warp owns TMEM lane quarter q (32 tile rows).  Per 32-column chunk: the accumulator row a lane
// holds goes to smem as 8 STS.128 (row stride 36 floats: conflict-free per 8-lane phase), comes back
// as LDS.128 with 8 lanes covering one row, and leaves as STG.128: 4 full 128-byte lines per store.
template <int ACT, bool HAS_RES, bool UP2, bool DBG>
__device__ __forceinline__ void epilogue_loop(const Params& p, float* stg, float* s_bias, uint32_t tmem_base,
                                              uint32_t bar0, int item0, int item1, int q, int lane, int ewarp, int grp) {
    const int sub_r = lane >> 3;        // row within a 4-row group
    const int sub_c = (lane & 7) << 2;  // first of this lane's 4 columns within the chunk
    uint32_t it = 0;
    Ring racc;
    int nt = item0 / p.m_tiles, mt = item0 - nt * p.m_tiles;
    int bias_nt = -1;
    for (int item = item0; item < item1; ++item, ++it) {
        if (p.nEpi == 2 && (int)(it & 1u) != grp) {  // the other epilogue group drains this tile
            racc.advance(p.nAcc);
            if (++mt == p.m_tiles) { mt = 0; ++nt; }
            continue;
        }
        if (nt != bias_nt) {
            // the n tile's bias lives in shared memory: a global load per 32-column chunk would put an L2
            // round trip on the critical path of every chunk (named barrier 1+grp = the 4 warps of this epilogue group)
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            for (int i = ewarp * 32 + lane; i < p.BN; i += 128) {
                const int n = nt * p.BN + i;
                s_bias[i] = n < p.N ? __ldg(p.bias + n) : 0.0f;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            bias_nt = nt;
        }
        const uint32_t acc = racc.slot;
        const int row0 = mt * BM + q * 32;
        const int ncols = min(p.BN, p.N - nt * p.BN);  // valid columns of this n tile (multiple of 8)
        const int rows = min(32, p.M - row0);          // valid rows of this warp's slab (may be <= 0)
        const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.acc_stride;
        bool waited = false;
        float v[32];
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            const int c = c0 + sub_c;
            const int n = nt * p.BN + c;
            const bool col_ok = c < ncols;
            float4 rv[8];
            if (HAS_RES) {  // residual rows of this chunk: issued before the accumulator wait so DRAM latency overlaps
                const float* rp = p.res + (size_t)(row0 + sub_r) * p.ld_res + n;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    rv[i] = (col_ok && sub_r + 4 * i < rows) ? ldg4(rp + (size_t)(4 * i) * p.ld_res) : make_float4(0.f, 0.f, 0.f, 0.f);
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
                *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            // v now lives in shared memory: fetch the next chunk into the same registers while this one is stored
            if (!HAS_RES && c0 + 32 < ncols) tmem_ld32_issue(tsrc + c0 + 32, v);
            __syncwarp();
            if (col_ok) {
                const float4 bv = *reinterpret_cast<const float4*>(s_bias + c);
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
        if (!waited) {  // no valid columns (cannot happen for a well-formed tiling); keep the protocol in step
            mbar_wait(bar0 + 8u * (BAR_ACC_FULL + acc), racc.phase, 6);
            tc_fence_after();
        }
        tc_fence_before();
        mbar_arrive(bar0 + 8u * (BAR_ACC_EMPTY + acc));
        if (ewarp == 0 && lane == 0) dbg_mark(p, 7, it);
        racc.advance(p.nAcc);
        if (++mt == p.m_tiles) { mt = 0; ++nt; }
    }
}

// ---- converters: raw fp32 A tile -> (hi, lo) TF32 tiles; the SE gate is folded in here -------------
// 8 warps; each thread owns 4 of the tile's 1024 16-byte chunks.  hi overwrites the raw tile in place
// (an elementwise map keeps the TMA swizzle), lo goes to a slot of the short lo ring.
template <bool HAS_SCALE, bool DBG>
__device__ __forceinline__ void converter_loop(const Params& p, uint8_t* gbase, uint32_t a_off, uint32_t l_off,
                                               uint32_t bar0, int item0, int item1, int ct, int grp) {
    Ring ra, rl;
    uint32_t dq = 0;
    int mt0 = item0 % p.m_tiles;
    for (int item = item0; item < item1;) {
        const int g = group_size(p, item, item1, mt0);
        for (int kb = 0; kb < p.KB; ++kb)
        for (int j = 0; j < g; ++j, ++dq) {
            const int mt = mt0 + j;
            if ((int)(dq & 1u) == grp) {
                mbar_wait(bar0 + 8u * (BAR_A_FULL + ra.slot), ra.phase, 5);
                if (ct == 0) dbg_mark(p, 1, dq);
                mbar_wait(bar0 + 8u * (BAR_L_EMPTY + rl.slot), rl.phase ^ 1u, 7);
                if (ct == 0) dbg_mark(p, 2, dq);
                float4* hi = reinterpret_cast<float4*>(gbase + a_off + ra.slot * (uint32_t)A_TILE_BYTES);
                float4* lo = reinterpret_cast<float4*>(gbase + l_off + rl.slot * (uint32_t)A_TILE_BYTES);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = hi[ct + (half * 4 + j) * 128];
                    if (HAS_SCALE) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = ct + (half * 4 + j) * 128;
                            const int r = i >> 3;
                            const int k = kb * BK + (((i & 7) ^ (r & 7)) << 2);  // undo the 128B swizzle
                            const int row = mt * BM + r;
                            if (k < p.K && row < p.M) {
                                const float4 g = ldg4(p.scale + (size_t)(row / p.rows_per_img) * p.K + k);
                                v[j].x *= g.x; v[j].y *= g.y; v[j].z *= g.z; v[j].w *= g.w;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 h, l;
                        h.x = tf32_rna(v[j].x); h.y = tf32_rna(v[j].y); h.z = tf32_rna(v[j].z); h.w = tf32_rna(v[j].w);
                        l.x = tf32_lo(v[j].x, h.x); l.y = tf32_lo(v[j].y, h.y);
                        l.z = tf32_lo(v[j].z, h.z); l.w = tf32_lo(v[j].w, h.w);
                        hi[ct + (half * 4 + j) * 128] = h;
                        lo[ct + (half * 4 + j) * 128] = l;
                    }
                }
                fence_proxy_async();
                mbar_arrive(bar0 + 8u * (BAR_A_CONV + ra.slot));
                if (ct == 0) dbg_mark(p, 3, dq);
            }
            ra.advance(p.nA);
            rl.advance(p.nL);
        }
        item += g;
        mt0 += g;
        if (mt0 == p.m_tiles) mt0 = 0;
    }
}

// ---- the GEMM ------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A ring: nA x 16K raw->hi][lo ring: nL x 16K][B slots: (hi | lo) x nB][epilogue transpose][barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_slot_bytes = 2u * p.BN * 128u;
    const uint32_t a_off = 0, l_off = a_off + p.nA * (uint32_t)A_TILE_BYTES;
    const uint32_t b_off = l_off + p.nL * (uint32_t)A_TILE_BYTES;
    const uint32_t epi_off = b_off + p.nB * b_slot_bytes;
    const uint32_t bar_off = epi_off + p.nEpi * (uint32_t)EPI_GROUP_BYTES;
    const uint32_t bar0 = base + bar_off;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8u * BAR_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nA; ++s) {
            mbar_init(bar0 + 8u * (BAR_A_FULL + s), 1);
            mbar_init(bar0 + 8u * (BAR_A_CONV + s), NUM_CONVERTERS / 2);
            mbar_init(bar0 + 8u * (BAR_A_EMPTY + s), 1);
        }
        for (int s = 0; s < p.nL; ++s) mbar_init(bar0 + 8u * (BAR_L_EMPTY + s), 1);
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

    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail; from here on the
    // kernel reads what that kernel wrote
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            Ring ra, rb;
            bool b_loaded = false;
            uint32_t dq = 0;
            int nt = item0 / p.m_tiles, mt = item0 - nt * p.m_tiles;
            for (int item = item0; item < item1;) {
                const int g = group_size(p, item, item1, mt);
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)nt * p.KB * b_slot_bytes;
                for (int kb = 0; kb < p.KB; ++kb) {
                    if (!(p.resident && b_loaded)) {
                        mbar_wait(bar0 + 8u * (BAR_B_EMPTY + rb.slot), rb.phase ^ 1u, 0);
                        mbar_expect_tx(bar0 + 8u * (BAR_B_FULL + rb.slot), b_slot_bytes);
                        bulk_load(base + b_off + rb.slot * b_slot_bytes, wsrc + (size_t)kb * b_slot_bytes, b_slot_bytes,
                                  bar0 + 8u * (BAR_B_FULL + rb.slot));
                        rb.advance(p.nB);
                    }
                    for (int j = 0; j < g; ++j) {
                        mbar_wait(bar0 + 8u * (BAR_A_EMPTY + ra.slot), ra.phase ^ 1u, 1);
                        mbar_expect_tx(bar0 + 8u * (BAR_A_FULL + ra.slot), A_TILE_BYTES);
                        tma_load_2d(base + a_off + ra.slot * (uint32_t)A_TILE_BYTES, &tmA,
                                    bar0 + 8u * (BAR_A_FULL + ra.slot), kb * BK, (mt + j) * BM);
                        dbg_mark(p, 0, dq++);
                        ra.advance(p.nA);
                    }
                }
                b_loaded = true;
                item += g;
                mt += g;
                if (mt == p.m_tiles) { mt = 0; ++nt; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop converged (addresses stay in uniform registers);
        // one elected lane issues the tcgen05 instructions =====
        Ring ra, rl, rb, racc;
        uint32_t dq = 0, b_seen = 0;
        // descriptors differ only in the 14-bit start-address field: build one, then add (bytes >> 4)
        const uint64_t desc0 = make_desc_sw128(base);
        const int k_tail = (p.K - (p.KB - 1) * BK + 7) / 8;  // 8-wide K steps of the last k-block
        int mt = item0 % p.m_tiles;
        for (int item = item0; item < item1;) {
            const int g = group_size(p, item, item1, mt);
            // accumulators of the group's tiles: consecutive slots of the ring
            Ring racc1 = racc;
            racc1.advance(p.nAcc);
            for (int kb = 0; kb < p.KB; ++kb) {
                uint32_t slot;
                if (p.resident) {
                    slot = kb;
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
                const int ksteps = (kb == p.KB - 1) ? k_tail : BK / 8;  // skip all-zero K steps of the tail
                for (int j = 0; j < g; ++j) {
                    const Ring& rc = j ? racc1 : racc;
                    if (kb == 0) {  // first MMA into this tile's accumulator: the epilogue must have drained it
                        mbar_wait(bar0 + 8u * (BAR_ACC_EMPTY + rc.slot), rc.phase ^ 1u, 2);
                    }
                    mbar_wait(bar0 + 8u * (BAR_A_CONV + ra.slot), ra.phase, 4);
                    tc_fence_after();
                    if (lane == 0) dbg_mark(p, 4, dq);
                    const uint32_t d_tmem = tmem_base + rc.slot * (uint32_t)p.acc_stride;
                    const uint64_t dah = desc0 + ((a_off + ra.slot * (uint32_t)A_TILE_BYTES) >> 4);
                    const uint64_t dal = desc0 + ((l_off + rl.slot * (uint32_t)A_TILE_BYTES) >> 4);
                    if (elect_one()) {
                        for (int k8 = 0; k8 < ksteps; ++k8) {
                            const uint64_t ko = (uint64_t)(k8 * 2);  // 32 bytes per 8-wide TF32 K step
                            umma_tf32(d_tmem, dal + ko, dbh + ko, p.idesc, (kb | k8) ? 1u : 0u);  // small terms first
                            umma_tf32(d_tmem, dah + ko, dbl + ko, p.idesc, 1u);
                            umma_tf32(d_tmem, dah + ko, dbh + ko, p.idesc, 1u);
                        }
                        umma_commit(bar0 + 8u * (BAR_A_EMPTY + ra.slot));
                        umma_commit(bar0 + 8u * (BAR_L_EMPTY + rl.slot));
                        if (!p.resident && j == g - 1) umma_commit(bar0 + 8u * (BAR_B_EMPTY + slot));
                        if (kb == p.KB - 1) umma_commit(bar0 + 8u * (BAR_ACC_FULL + rc.slot));
                    }
                    __syncwarp();
                    if (lane == 0) dbg_mark(p, 5, dq);
                    ++dq;
                    ra.advance(p.nA);
                    rl.advance(p.nL);
                }
            }
            racc.advance(p.nAcc);
            if (g == 2) racc.advance(p.nAcc);
            item += g;
            mt += g;
            if (mt == p.m_tiles) mt = 0;
        }
    } else if (warp < 2 + NUM_CONVERTERS / 32) {
        const int ct = (threadIdx.x - 64) & 127, cg = (threadIdx.x - 64) >> 7;  // thread within group, group
        if (p.scale != nullptr) converter_loop<true, DBG>(p, gbase, a_off, l_off, bar0, item0, item1, ct, cg);
        else converter_loop<false, DBG>(p, gbase, a_off, l_off, bar0, item0, item1, ct, cg);
    } else {
        // ===== epilogue =====
        const int ew8 = warp - (2 + NUM_CONVERTERS / 32);  // 0..7
        const int ew = ew8 & 3, eg = ew8 >> 2;             // warp within its group, group
        float* stg = reinterpret_cast<float*>(gbase + epi_off + eg * EPI_GROUP_BYTES) + ew * 32 * EPI_LD;
        float* s_bias = reinterpret_cast<float*>(gbase + epi_off + eg * EPI_GROUP_BYTES + 4 * 32 * EPI_LD * 4);
        if (eg < p.nEpi) {
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
    int BN, n_tiles, KB, nA, nL, nB, nAcc, nEpi, resident, tmem_cols, acc_stride;
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
    const long long tile = A_TILE_BYTES;
    bool ok = false;
    long long fixed = 0;
    // Resident weights (the common case for the huge-M layers): two epilogue groups, 4 lo slots, the rest
    // of shared memory is the raw A ring.  Streamed weights: one epilogue group and 2 lo slots leave room
    // for one more weight slot, which is what the MMA waits on there.
    {
        t.nEpi = 2;
        t.nL = 4;
        fixed = 1024 /*alignment slack*/ + (long long)t.nEpi * EPI_GROUP_BYTES + BAR_BYTES;
        const long long avail = SMEM_LIMIT - fixed;
        if (t.n_tiles == 1 && t.KB <= MAX_B_SLOTS && t.KB * slot + (4 + t.nL) * tile <= avail) {
            t.resident = 1;
            t.nB = t.KB;
            long long na = (avail - t.nB * slot - t.nL * tile) / tile;
            // even ring: each slot then always belongs to the same converter group, which sees every phase of the
            // slot's mbarrier (an odd ring makes a group skip the phase the other group consumes; see pwconv_ts.cu)
            t.nA = (int)(na > MAX_A_STAGES ? MAX_A_STAGES : na) & ~1;
            ok = true;
        }
    }
    if (!ok) {
        t.nEpi = 1;
        t.nL = 2;
        t.resident = 0;
        fixed = 1024 + (long long)t.nEpi * EPI_GROUP_BYTES + BAR_BYTES;
        const long long avail = SMEM_LIMIT - fixed;
        long long nb = (avail - (t.nL + 4) * tile) / slot;  // keep 4 raw A slots when the weights leave room
        if (nb > 4) nb = 4;
        if (nb > t.KB) nb = t.KB;
        if (nb < 2) nb = 2;
        t.nB = (int)nb;
        if (t.nB >= t.KB && t.n_tiles == 1 && t.KB <= MAX_B_SLOTS) t.resident = 1;
        long long na = (avail - t.nB * slot - t.nL * tile) / tile;
        if (na > MAX_A_STAGES) na = MAX_A_STAGES;
        na &= ~1ll;
        t.nA = (int)na;
        ok = na >= 2;
    }
    if (!ok) return false;
    t.acc_stride = (t.BN + 31) / 32 * 32;  // the epilogue reads TMEM in 32-column chunks
    t.nAcc = 512 / t.acc_stride;  // a deep accumulator ring hides the MMA -> epilogue -> MMA round trip
    if (t.nAcc > MAX_ACC) t.nAcc = MAX_ACC;
    if (t.nAcc < 2) return false;
    int cols = 32;
    while (cols < t.nAcc * t.acc_stride) cols *= 2;
    if (cols > 512) return false;
    t.tmem_cols = cols;
    t.smem = (size_t)(fixed + (t.nA + t.nL) * tile + t.nB * slot);
    return t.smem <= (size_t)SMEM_LIMIT;
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
    YR_CHECK_ARG((op.Ho == op.H && op.Wo == op.W) || (op.Ho == 2 * op.H && op.Wo == 2 * op.W && !op.res),
                 "pw_tc: output must be HxW, or 2Hx2W (fused nearest upsampling, no residual): got %dx%d for %dx%d", op.Ho, op.Wo,
                 op.H, op.W);
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
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tc::l2_promotion(),
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
    p.img_w = op.W;
    p.up2 = (op.Ho == 2 * op.H && op.Wo == 2 * op.W) ? 1 : 0;
    p.act = op.act;
    p.nA = t.nA;
    p.nL = t.nL;
    p.nAcc = t.nAcc;
    p.nEpi = t.nEpi;
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
    static DeviceOnce attr_once;  // function attributes are per device
    bool& attr_set = attr_once.cur();
    if (!attr_set) {
        if (cudaFuncSetAttribute(tc::pw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT) !=
                cudaSuccess ||
            cudaFuncSetAttribute(tc::pw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT) !=
                cudaSuccess) {
            set_error("pw_tc: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return YR_ERR_CUDA;
        }
        attr_set = true;
    }
    p.dbg = nullptr;
    static const bool debug = getenv("YR_PW_TC_DEBUG") != nullptr;  // developer aid only: timeline of CTA 0
    if (debug) {
        static long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc(&dbuf, 8 * tc::DBG_EV * sizeof(long long));
        cudaMemsetAsync(dbuf, 0, 8 * tc::DBG_EV * sizeof(long long), s);
        p.dbg = dbuf;
    }
    if (launch_pdl(debug ? tc::pw_tc_kernel<true> : tc::pw_tc_kernel<false>, dim3(grid), dim3(tc::NUM_THREADS), t.smem, s, tm,
                   p) != cudaSuccess) {
        set_error("pw_tc: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return YR_ERR_CUDA;
    }
    if (debug) {
        static long long h[8 * tc::DBG_EV];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[8] = {"tma_issue", "a_full_seen", "lo_empty_seen", "conv_done", "mma_start", "mma_issued",
                                "acc_full_seen", "epi_done"};
        long long t0 = h[0];
        fprintf(stderr, "pw_tc timeline K=%d N=%d KB=%d nA=%d nL=%d nB=%d resident=%d items/cta=%d (cycles since first TMA)\n",
                p.K, p.N, p.KB, p.nA, p.nL, p.nB, p.resident, p.items_per_cta);
        for (int r = 0; r < 8; ++r) {
            fprintf(stderr, "%-14s", names[r]);
            for (int i = 0; i < 40 && h[r * tc::DBG_EV + i]; ++i) fprintf(stderr, " %6lld", h[r * tc::DBG_EV + i] - t0);
            fprintf(stderr, "\n");
        }
    }
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
