// Optimizer step of the reference's training loop on a flat fp32 parameter shard.
//
// Reference: tf.keras.optimizers.Adam(lr, epsilon=1e-8) (code/train.py:158-160,195-197), whose dense update is TF's
// ApplyAdam functor (third-party, un-vendored, unpinned - published algorithm restated; oracle/optim.py):
//     alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//     m += (g - m) * (1 - beta1);   v += (g*g - v) * (1 - beta2);   var -= (m * alpha) / (sqrt(v) + epsilon)
// One pass over the shard: 16 bytes read + 12 written per parameter, float4 vectorised - an HBM-bound stream.
// In the data-parallel step it runs on the rank's 1/N shard between the gradient reduce-scatter and the parameter
// all-gather (yoloret_b200/parallel.py: GradBucket, ShardedAdam).
#include "yr_common.cuh"

namespace yr {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float alpha, float omb1, float omb2, float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
            long long n, float alpha, float omb1, float omb2, float eps) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4*>(param)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_one(p.x, g.x, mm.x, vv.x, alpha, omb1, omb2, eps);
        adam_one(p.y, g.y, mm.y, vv.y, alpha, omb1, omb2, eps);
        adam_one(p.z, g.z, mm.z, vv.z, alpha, omb1, omb2, eps);
        adam_one(p.w, g.w, mm.w, vv.w, alpha, omb1, omb2, eps);
        reinterpret_cast<float4*>(param)[i] = p;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail
        const long long i = (n4 << 2) + threadIdx.x;
        float p = param[i], mm = m[i], vv = v[i];
        adam_one(p, grad[i], mm, vv, alpha, omb1, omb2, eps);
        param[i] = p; m[i] = mm; v[i] = vv;
    }
}

}  // namespace yr

using namespace yr;

extern "C" int yr_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                            float beta2, float epsilon, int64_t step, void* stream) {
    YR_CHECK_ARG(param && grad && m && v, "adam: null pointer");
    YR_CHECK_ARG(n >= 0 && step >= 1, "adam: n >= 0 and step >= 1 (1-based like Keras' iterations + 1)");
    YR_CHECK_ARG(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "adam: pointers must be 16-byte aligned");
    if (n == 0) return YR_OK;
    // float32 like the TF variables: beta^t by powf, alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    const float b1p = powf(beta1, (float)step), b2p = powf(beta2, (float)step);
    const float alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, (long long)n, alpha, 1.0f - beta1,
                                                                    1.0f - beta2, epsilon);
    YR_CHECK_LAUNCH("adam");
    return YR_OK;
}
