// Instantiation unit of the tcgen05 conv kernel: PLANES = 2, N tile = 128 (see conv_umma_kernel.cuh).
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(128, 2, FISR_FOR_EPI_TRAIN)
}  // namespace convk
}  // namespace fisr
