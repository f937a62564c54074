// at_fused_imma.cu -- fused localization kernel, int8 tensor-core form (AT_KERNEL_IMMA).
// Placeholder until the byte-split Toeplitz x Hankel kernel lands: no shape is supported, so
// AT_KERNEL_AUTO resolves to the integer-pipe kernel.
#include "at_internal.h"

bool at_fused_imma_supports(const AtShape &) { return false; }

cudaError_t at_launch_fused_imma(const AtShape &, const AtFusedParams &, int, cudaStream_t)
{
    return cudaErrorInvalidValue;
}
