// evaluator.cu — policy/value evaluation of the gathered leaf batch.
//
// Replaces the tensor part of Model::predict (ref: src/model/mod.rs:60-67,95) and Net::forward
// (ref: src/model/connect_four.rs:75-81; architecture model/mod.rs:152-184, connect_four.rs:50-73):
//   encode (3 planes) -> stem conv3x3+BN+ReLU -> 4 x [conv-BN-ReLU-conv-BN, +skip, ReLU]
//   -> policy head conv3x3(64->32)+BN+ReLU -> Linear -> softmax
//   -> value  head conv3x3(64->3)+BN+ReLU  -> Linear -> tanh
// BatchNorm (eval) is folded into the convs on the host; weights are bf16, accumulation f32,
// activations are rounded to bf16 between layers (both kernels do the same roundings).
//
// Two kernels:
//   k_eval_umma  — the product path: implicit-GEMM 3x3 convs on tcgen05 tensor cores, accumulators
//                  in TMEM, weights streamed by TMA bulk copies (see evaluator_umma.cuh).
//   k_eval_simt  — CUDA-core cross-check of the same arithmetic (SPB_FLAG_EVAL_SIMT), used by tests
//                  to localise a disagreement; never selected implicitly.
#include <cuda_bf16.h>

#include <cstring>
#include <vector>

#include "evaluator.cuh"
#include "evaluator_umma.cuh"

namespace spb {

namespace {

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

inline uint16_t host_f2bf(float f) {   // round-to-nearest-even, NaN-safe enough for finite weights
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// 3x3 conv, pad 1, over one board held in shared memory: in[IC][P] -> f(oc, p, acc + bias).
template <class G, class F>
__device__ __forceinline__ void conv3x3_simt(const float* in, int IC, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                                             int OC, F&& f) {
  constexpr int R = G::ROWS, C = G::COLS, P = R * C;
  for (int idx = threadIdx.x; idx < OC * P; idx += blockDim.x) {
    const int oc = idx / P, p = idx % P, r = p / C, c = p % C;
    const uint16_t* wr = w + (size_t)oc * IC * 9;
    float acc = 0.0f;
    for (int ic = 0; ic < IC; ++ic) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int rr = r + ky - 1;
        if (rr < 0 || rr >= R) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int cc = c + kx - 1;
          if (cc < 0 || cc >= C) continue;
          acc = fmaf(bf16_bits_to_float(wr[ic * 9 + ky * 3 + kx]), in[ic * P + rr * C + cc], acc);
        }
      }
    }
    f(oc, p, acc + bias[oc]);
  }
}

template <class G>
__global__ void __launch_bounds__(256) k_eval_simt(Evaluator::DevNet net, const PState* __restrict__ states,
                                                   const uint32_t* __restrict__ list, const uint32_t* __restrict__ count,
                                                   uint32_t max_n, float* out, int stride, float* logits_out) {
  constexpr int R = G::ROWS, C = G::COLS, P = R * C, A = G::A;
  __shared__ float xa[NET_HIDDEN * P];
  __shared__ float ha[NET_HIDDEN * P];
  __shared__ float head[(NET_POLICY_CH + NET_VALUE_CH) * P];
  __shared__ float logits[A + 1];
  const uint32_t n = min(*count, max_n);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t slot = list ? list[i] : i;
    const PState st = states[slot];
    for (int idx = threadIdx.x; idx < 3 * P; idx += blockDim.x) {
      int plane = idx / P, p = idx % P;
      ha[idx] = G::encode_cell(st, plane, p / C, p % C);                       // get_encoding
    }
    __syncthreads();
    conv3x3_simt<G>(ha, 3, net.w_simt[0], net.bias[0], NET_HIDDEN,
                    [&](int oc, int p, float v) { xa[oc * P + p] = round_bf16(fmaxf(v, 0.0f)); });   // stem
    __syncthreads();
    for (int b = 0; b < NET_BLOCKS; ++b) {                                     // resnet_block, model/mod.rs:152-165
      conv3x3_simt<G>(xa, NET_HIDDEN, net.w_simt[1 + 2 * b], net.bias[1 + 2 * b], NET_HIDDEN,
                      [&](int oc, int p, float v) { ha[oc * P + p] = round_bf16(fmaxf(v, 0.0f)); });
      __syncthreads();
      conv3x3_simt<G>(ha, NET_HIDDEN, net.w_simt[2 + 2 * b], net.bias[2 + 2 * b], NET_HIDDEN,
                      [&](int oc, int p, float v) { xa[oc * P + p] = round_bf16(fmaxf(v + xa[oc * P + p], 0.0f)); });   // (x + f(x)).relu()
      __syncthreads();
    }
    conv3x3_simt<G>(xa, NET_HIDDEN, net.w_simt[9], net.bias[9], NET_POLICY_CH,
                    [&](int oc, int p, float v) { head[oc * P + p] = fmaxf(v, 0.0f); });
    conv3x3_simt<G>(xa, NET_HIDDEN, net.w_simt[10], net.bias[10], NET_VALUE_CH,
                    [&](int oc, int p, float v) { head[(NET_POLICY_CH + oc) * P + p] = fmaxf(v, 0.0f); });
    __syncthreads();
    // Linear layers: one warp per output (flat_view index = channel*P + row*C + col).
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < A + 1; o += (int)(blockDim.x >> 5)) {
      const bool is_value = o == A;
      const int K = is_value ? NET_VALUE_CH * P : NET_POLICY_CH * P;
      const float* wrow = is_value ? net.vfc_w : net.pfc_w + (size_t)o * K;
      const float* src = is_value ? head + NET_POLICY_CH * P : head;
      float acc = 0.0f;
      for (int k = lane; k < K; k += 32) acc = fmaf(wrow[k], src[k], acc);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) logits[o] = acc + (is_value ? net.vfc_b[0] : net.pfc_b[o]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = logits[0];
      for (int a = 1; a < A; ++a) m = fmaxf(m, logits[a]);
      float e[A], s = 0.0f;
      for (int a = 0; a < A; ++a) { e[a] = expf(logits[a] - m); s += e[a]; }
      float* o = out + (size_t)slot * stride;
      for (int a = 0; a < A; ++a) o[a] = e[a] / s;                             // softmax, model/mod.rs:63
      o[A] = tanhf(logits[A]);                                                 // connect_four.rs:71
      if (logits_out) for (int a = 0; a < A; ++a) logits_out[(size_t)slot * A + a] = logits[a];
    }
    __syncthreads();
  }
}

}  // namespace

Evaluator::~Evaluator() {
  if (d_blob_) cudaFree(d_blob_);
}

double Evaluator::flops_per_position() const {
  const double P = (double)rows_ * cols_;
  double macs = P * 9.0 * (3.0 * NET_HIDDEN + 2.0 * NET_BLOCKS * NET_HIDDEN * NET_HIDDEN + NET_HIDDEN * (NET_POLICY_CH + NET_VALUE_CH));
  macs += NET_POLICY_CH * P * actions_ + NET_VALUE_CH * P;
  return 2.0 * macs;
}

bool Evaluator::upload(const HostNet& net, std::string* err) {
  const int P = net.rows * net.cols, A = net.actions;
  // blob layout (all sections 256-byte aligned)
  std::vector<uint8_t> blob;
  auto reserve = [&](size_t bytes) { size_t off = (blob.size() + 255) & ~(size_t)255; blob.resize(off + bytes, 0); return off; };
  size_t off_wsimt[NET_CONVS], off_bias[NET_CONVS];
  for (int i = 0; i < NET_CONVS; ++i) {
    const auto& c = net.conv[i];
    off_wsimt[i] = reserve((size_t)c.oc * c.ic * 9 * 2);
    uint16_t* w = reinterpret_cast<uint16_t*>(blob.data() + off_wsimt[i]);
    for (size_t k = 0; k < c.w.size(); ++k) w[k] = host_f2bf(c.w[k]);
    off_bias[i] = reserve((size_t)c.oc * 4);
    std::memcpy(blob.data() + off_bias[i], c.b.data(), (size_t)c.oc * 4);
  }
  const size_t off_pw = reserve(net.pfc_w.size() * 4), off_pb = reserve(net.pfc_b.size() * 4);
  const size_t off_vw = reserve(net.vfc_w.size() * 4), off_vb = reserve(net.vfc_b.size() * 4);
  std::memcpy(blob.data() + off_pw, net.pfc_w.data(), net.pfc_w.size() * 4);
  std::memcpy(blob.data() + off_pb, net.pfc_b.data(), net.pfc_b.size() * 4);
  std::memcpy(blob.data() + off_vw, net.vfc_w.data(), net.vfc_w.size() * 4);
  std::memcpy(blob.data() + off_vb, net.vfc_b.data(), net.vfc_b.size() * 4);
  // tcgen05 operand images + epilogue tables
  std::vector<uint8_t> umma_img;
  umma::pack_weights(net, &umma_img);
  const size_t off_umma = reserve(umma_img.size());
  std::memcpy(blob.data() + off_umma, umma_img.data(), umma_img.size());

  // new image first, then swap: a failed upload keeps the previous checkpoint usable
  void* nblob = nullptr;
  cudaError_t e = cudaMalloc(&nblob, blob.size());
  if (e == cudaSuccess) e = cudaMemcpy(nblob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (nblob) cudaFree(nblob);
    cudaGetLastError();
    *err = std::string("weight upload: ") + cudaGetErrorString(e);
    return false;
  }
  if (d_blob_) cudaFree(d_blob_);
  d_blob_ = nblob;
  blob_bytes_ = blob.size();
  const uint8_t* d = static_cast<const uint8_t*>(d_blob_);
  for (int i = 0; i < NET_CONVS; ++i) {
    dev_.w_simt[i] = reinterpret_cast<const uint16_t*>(d + off_wsimt[i]);
    dev_.bias[i] = reinterpret_cast<const float*>(d + off_bias[i]);
  }
  dev_.pfc_w = reinterpret_cast<const float*>(d + off_pw);
  dev_.pfc_b = reinterpret_cast<const float*>(d + off_pb);
  dev_.vfc_w = reinterpret_cast<const float*>(d + off_vw);
  dev_.vfc_b = reinterpret_cast<const float*>(d + off_vb);
  dev_.w_umma = reinterpret_cast<const uint16_t*>(d + off_umma);
  dev_.rows = net.rows; dev_.cols = net.cols; dev_.actions = A;
  game_ = net.game; rows_ = net.rows; cols_ = net.cols; actions_ = A;
  (void)P;
  loaded_ = true;
  return true;
}

cudaError_t Evaluator::launch(const PState* states, const uint32_t* list, const uint32_t* count_dev, uint32_t max_n, float* out,
                              int stride, float* logits_out, bool simt, cudaStream_t stream, bool overlap) {
  if (!loaded_) return cudaErrorNotReady;
  if (simt) {
    const unsigned grid = std::min<unsigned>(max_n, 148u * 4u);
    if (game_ == SPB_GAME_CONNECT4)
      k_eval_simt<Connect4><<<grid, 256, 0, stream>>>(dev_, states, list, count_dev, max_n, out, stride, logits_out);
    else
      k_eval_simt<TicTacToe><<<grid, 256, 0, stream>>>(dev_, states, list, count_dev, max_n, out, stride, logits_out);
    return cudaGetLastError();
  }
  return umma::launch(dev_, game_, states, list, count_dev, max_n, out, stride, logits_out, stream, overlap);
}

cudaError_t Evaluator::launch_ring(const Trees& T, const AsyncCtl& C, cudaStream_t stream) {
  if (!loaded_) return cudaErrorNotReady;
  return umma::launch_ring(dev_, game_, T, C, stream);
}

}  // namespace spb
