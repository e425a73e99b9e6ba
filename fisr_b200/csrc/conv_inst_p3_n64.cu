// Instantiation unit of the tcgen05 conv kernel: PLANES = 3 (f16f8: fp16 main term + fp8 cross terms), N tile = 64
// (see conv_umma_kernel.cuh).  Inference-only epilogues: training keeps the f16x3 operand format.
#include "conv_umma_kernel.cuh"

namespace fisr {
namespace convk {
FISR_CONV_FAMILY(64, 3, FISR_FOR_EPI)
}  // namespace convk
}  // namespace fisr
