// Instantiation unit of the tcgen05 conv kernel: PLANES = 3 (f16f8), N tile = 32: the FI-SR conv/2 head evaluated at input
// resolution with its depth_to_space folded into the weights (4 x 6 = 24 output columns), see fisr_api.cu.
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(32, 3, FISR_FOR_EPI_NARROW)
}  // namespace convk
}  // namespace fisr
