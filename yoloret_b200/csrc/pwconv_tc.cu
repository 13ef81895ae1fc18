// tcgen05 3xTF32 pointwise convolution (placeholder until the kernel lands).
#include "yr_common.cuh"
namespace yr {
int launch_pw_tc(const yr_op& op, cudaStream_t s) {
    (void)op; (void)s;
    set_error("pw: tcgen05 variant not built yet");
    return YR_ERR_UNSUPPORTED;
}
}  // namespace yr
