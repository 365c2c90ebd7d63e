// evaluator_umma.cuh — interface of the tcgen05 evaluator kernel (evaluator_umma.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "evaluator.cuh"

namespace spb {
namespace umma {

// Packs the folded net into the device image the tcgen05 kernel streams with TMA bulk copies.
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);

cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream);

}  // namespace umma

namespace umma_v1 {   // first version (one MMA group per tap, N = 64); cross-check only
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);
cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream);
}  // namespace umma_v1

namespace umma_v2 {   // kx-pair variant: centre + right taps share one A fetch (N = 128), see evaluator_umma_v2.cu
void pack_weights(const HostNet& net, std::vector<uint8_t>* out);
cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list,
                   const uint32_t* count_dev, uint32_t max_n, float* out, int stride, float* logits_out,
                   cudaStream_t stream);
}  // namespace umma_v2
}  // namespace spb
