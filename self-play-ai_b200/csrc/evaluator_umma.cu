// evaluator_umma.cu — placeholder until the hardware probe (tools/umma_probe.cu) has confirmed the
// descriptor layout; replaced by the tcgen05 kernel.
#include "evaluator_umma.cuh"

namespace spb {
namespace umma {

void pack_weights(const HostNet&, std::vector<uint8_t>* out) { out->assign(256, 0); }

cudaError_t launch(const Evaluator::DevNet&, int, const PState*, const uint32_t*, const uint32_t*, uint32_t, float*, int, float*,
                   cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace umma
}  // namespace spb
