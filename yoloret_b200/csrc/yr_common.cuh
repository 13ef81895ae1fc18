// Shared helpers for the yoloret_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/yoloret_b200.h"

namespace yr {

void set_error(const char* fmt, ...);

#define YR_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            yr::set_error(__VA_ARGS__);         \
            return YR_ERR_INVALID;              \
        }                                       \
    } while (0)

#define YR_CHECK_LAUNCH(what)                                                   \
    do {                                                                        \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess) {                                               \
            yr::set_error("%s: launch failed: %s", what, cudaGetErrorString(e__)); \
            return YR_ERR_CUDA;                                                 \
        }                                                                       \
    } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// A "done once" flag PER DEVICE: function attributes (the dynamic shared-memory limit) and the SM count belong to
// a device, so a process that drives several GPUs must set / query them once per device, not once per process.
struct DeviceOnce {
    bool done[64] = {};
    bool& cur() {
        int d = 0;
        cudaGetDevice(&d);
        return done[d & 63];
    }
};

// Activations of the graph: ReLU6 (tf.keras.layers.ReLU(6.)) and Swish
// (reference code/yolo3/efficientnet.py:327-331: x * sigmoid(x)).
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Swish in the conv epilogues: x * sigmoid(x) with the hardware exp2 / reciprocal (__expf, __fdividef: ~2 ulp each,
// i.e. ~3e-7 relative, far inside the 1e-3 output budget).  The IEEE expf + division version cost ~25 instructions
// per element and made the Swish depthwise layers of the detection heads issue-bound (72 us for a 27 us layer).
__device__ __forceinline__ float swish_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
    if (ACT == YR_ACT_RELU6) return fminf(fmaxf(v, 0.0f), 6.0f);
    if (ACT == YR_ACT_SWISH) return swish_fast(v);
    return v;
}

__device__ __forceinline__ float apply_act_rt(float v, int act) {
    if (act == YR_ACT_RELU6) return fminf(fmaxf(v, 0.0f), 6.0f);
    if (act == YR_ACT_SWISH) return swish_fast(v);
    return v;
}

// Programmatic dependent launch: every kernel of the step is launched with the programmatic-stream-serialization
// attribute, runs its prologue (barrier init, TMEM allocation, descriptor prefetch, weight loads - nothing the
// previous kernel writes), then waits here until the previous kernel in the stream has completed and its writes are
// visible.  The prologue and the launch latency overlap the previous kernel's tail; inside the captured CUDA graph
// the edges become programmatic dependencies.  Opt-in with YR_PDL=1 (measured on B200: 4.06 vs 4.03 ms per step, i.e.
// the step is the sum of its kernel bodies, not of launch gaps); default is plain stream order.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// acc += x * w on four lanes with two packed FFMA2 (fma.rn.f32x2: each lane is an IEEE fused multiply-add, so the bits
// equal four fmaf calls) - half the issue slots of scalar FFMAs in the issue-bound depthwise inner loops.
__device__ __forceinline__ void fma4(float4& acc, const float4& x, const float4& w) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(&acc);
    const unsigned long long* xx = reinterpret_cast<const unsigned long long*>(&x);
    const unsigned long long* ww = reinterpret_cast<const unsigned long long*>(&w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[0]) : "l"(xx[0]), "l"(ww[0]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[1]) : "l"(xx[1]), "l"(ww[1]));
}

// acc += a * b on two lanes with one packed FFMA2 (each lane an IEEE fma)
__device__ __forceinline__ void fma2(float2& acc, const float2& a, const float2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// per-op launchers (one .cu each)
int launch_stem(const yr_op& op, cudaStream_t s);
int launch_pw(const yr_op& op, cudaStream_t s);
int launch_pw_tc(const yr_op& op, cudaStream_t s);
int launch_pw_ts(const yr_op& op, cudaStream_t s);
int launch_pw_ts2(const yr_op& op, cudaStream_t s);
int launch_dwpw(const yr_op& op, cudaStream_t s);
int launch_dw(const yr_op& op, cudaStream_t s);
int launch_resample(const yr_op& op, cudaStream_t s);
int launch_rfcr(const yr_op& op, cudaStream_t s);
int launch_se(const yr_op& op, cudaStream_t s);
int launch_se_fc(const yr_op& op, cudaStream_t s);
int dw_se_slots(const yr_op& op);
bool dw_uses_tma(const yr_op& op);
int dw_tma_se_slots(const yr_op& op);
int launch_dw_tma(const yr_op& op, cudaStream_t s);

}  // namespace yr
