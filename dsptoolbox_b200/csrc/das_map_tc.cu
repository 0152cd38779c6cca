// Tensor-core (tcgen05) variants of the delay-and-sum map -- placeholder until the kernel lands.
#include "common.cuh"

int64_t b2_das_tc_operand_bytes(int, int, int) {
    b2_set_error("tensor-core map kernel not built");
    return B2_EUNSUPPORTED;
}
int b2_das_tc_prepare(const void *, int, int, int, int, int, void *, int, cudaStream_t) {
    b2_set_error("tensor-core map kernel not built");
    return B2_EUNSUPPORTED;
}
int b2_das_tc_map(const void *, int, int, const void *, const void *, double, const void *, const double *,
                  const void *, int64_t, int, int, void *, void *, int, cudaStream_t) {
    b2_set_error("tensor-core map kernel not built");
    return B2_EUNSUPPORTED;
}
