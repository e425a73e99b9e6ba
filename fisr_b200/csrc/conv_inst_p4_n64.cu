// Instantiation unit of the tcgen05 conv kernel: PLANES = 4 = f16f8 on CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256),
// N tile = 64 (see conv_umma_kernel.cuh).
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(64, 4, FISR_FOR_EPI)
}  // namespace convk
}  // namespace fisr
