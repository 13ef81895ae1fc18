// tcgen05 / TMA / mbarrier building blocks shared by the tensor-core kernels (pwconv_tc.cu, pwconv_ts.cu) and the
// TMA depthwise kernel (dwconv_tma.cu).
// Plain inline PTX for sm_100a; see DESIGN.md section 4 for how the kernels use them.
#pragma once
#include "yr_common.cuh"
#include <stdlib.h>
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time

namespace yr {
namespace tc {

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
    // the suspend-time hint lets the hardware park the warp until the phase completes, so idle roles do
    // not steal issue slots from working ones (measured: same or slightly better than the plain form)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (and fails the launch loudly) instead of hanging the GPU.  The bound is wall time
// (20 s on %globaltimer), not a spin count: under compute-sanitizer a legitimate wait can take millions of polls.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int what) {
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if ((spins & 255u) == 255u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) {
                printf("yoloret_b200 tcgen05 kernel: mbarrier wait timed out (role %d, block %d, thread %d)\n", what, blockIdx.x,
                       threadIdx.x);
                __trap();
            }
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive TMEM columns of this thread's lane -> registers.  The load is asynchronous: the registers are valid
// only after tmem_ld_wait(), so an epilogue can issue the load of its next chunk and hide the TMEM latency behind the
// stores of the current one.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    tmem_ld32_issue(taddr, v);
    tmem_ld_wait();
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

// The low word of the 3xTF32 split of an ACTIVATION: x - hi is exact in fp32 and is handed to the tensor core as is.
// kind::tf32 reads the top 19 bits of the 32-bit container and ignores the 13 low mantissa bits, so the operand is the
// truncation of x - hi: error <= 2^-10 |lo| <= 2^-21 |x| (rounding it first, as the weight packers do at no cost,
// would give 2^-22 |x|) against 2 integer-ALU instructions per element saved in every converter warp.  All three
// tensor-core kernels use this one helper, so they stay bit-identical to each other.
__device__ __forceinline__ float tf32_lo(float x, float hi) { return x - hi; }

struct Ring {  // (slot, phase) cursor of a circular buffer; no div/mod on the hot path
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++slot == n) { slot = 0; phase ^= 1u; }
    }
};

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
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

// L2 promotion of the activation tensor maps (experiment knob YR_TMA_L2PROMO = 0 none / 1 128 B / 2 256 B).
inline CUtensorMapL2promotion l2_promotion() {
    static const int v = [] {
        const char* e = getenv("YR_TMA_L2PROMO");
        return e ? atoi(e) : 1;
    }();
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

inline int num_sms() {  // of the CURRENT device (cached per device)
    static int cache[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& n = cache[dev & 63];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace tc
}  // namespace yr
