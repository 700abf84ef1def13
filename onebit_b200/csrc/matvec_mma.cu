// placeholder until the tensor-core variant lands
#include "common.cuh"
namespace onebit {
bool matvec_mma_supported(int64_t, int64_t, int64_t, int) { return false; }
int launch_matvec_mma(const void*, const int8_t*, const void*, const void*, float*, int64_t, int64_t, int64_t, int, int,
                      bool, cudaStream_t) {
    return fail(ONEBIT_ERR_INVALID_ARGUMENT, "mma variant not built");
}
}  // namespace onebit
