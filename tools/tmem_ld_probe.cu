// tmem_ld_probe.cu — TMEM -> register bandwidth probe (tcgen05.ld.32x32b.xN), not product code.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                 "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                 "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int X>
__global__ void probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    const uint32_t col = (uint32_t)((i * X) & 255);
    if (X == 8) tmem_ld8(base + col, r);
    if (X == 16) tmem_ld16(base + col, r);
    if (X == 32) tmem_ld32(base + col, r);
    tmem_wait();
#pragma unroll
    for (int j = 0; j < X; ++j) acc ^= r[j];
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
}

int main() {
  long long* d; uint32_t* s;
  CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&s, 64));
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    for (int X : {8, 16, 32}) {
      if (X == 8) probe<8><<<1, warps * 32>>>(iters, d, s);
      if (X == 16) probe<16><<<1, warps * 32>>>(iters, d, s);
      if (X == 32) probe<32><<<1, warps * 32>>>(iters, d, s);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
      double per = (double)c / iters;
      double bytes = (double)warps * 32 * X * 4;
      printf("warps %2d x%-2d: %.1f cycles per (ld+wait) per warp; SM-wide %.1f B/cycle\n", warps, X, per, bytes / per);
    }
  }
  return 0;
}
