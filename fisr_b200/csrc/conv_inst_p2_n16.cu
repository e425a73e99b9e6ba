// Instantiation unit of the tcgen05 conv kernel: PLANES = 2, N tile = 16 (see conv_umma_kernel.cuh).
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(16, 2, FISR_FOR_EPI_NARROW)
}  // namespace convk
}  // namespace fisr
