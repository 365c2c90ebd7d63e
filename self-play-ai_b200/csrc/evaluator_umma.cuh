// evaluator_umma.cuh — interface of the tcgen05 evaluator kernels (evaluator_umma_v2.cu: default; evaluator_umma_v1.cu: cross-check).
#pragma once
#include <cstdint>
#include <vector>

#include "evaluator.cuh"

namespace spb {
// pack_weights: packs the folded net into the device image the kernel streams with TMA bulk copies.
namespace umma_v1 {   // first version (one MMA group per tap, N = 64); cross-check only (SPB_FLAG_EVAL_V1)
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);
cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream);
}  // namespace umma_v1

namespace umma_v2 {   // DEFAULT: centre + right taps of a kernel row share one A fetch (N = 128), see evaluator_umma_v2.cu
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);
cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream, bool overlap);   // overlap: programmatic dependent launch (set-up may run under the previous kernel)
}  // namespace umma_v2

namespace umma_v3 {   // the kx-pair kernel on CTA pairs (cta_group::2, each CTA holds half of B), see evaluator_umma_v3.cu
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);
cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream, bool overlap);
}  // namespace umma_v3
}  // namespace spb
