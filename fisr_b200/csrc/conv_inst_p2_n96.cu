// Instantiation unit of the tcgen05 conv kernel: PLANES = 2, N tile = 96 (the 96-channel layers of PWC-Net, pwc_api.cu; activation,
// fp32-out and residual epilogues).
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(96, 2, FISR_FOR_EPI_PWC)
}  // namespace convk
}  // namespace fisr
