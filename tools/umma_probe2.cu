// umma_probe2.cu — hardware probe for the 2-CTA (cta_group::2) variant of the evaluator's MMA (not product code).
// Two CTAs of a cluster each hold 128 rows of A and one half (N/2 rows) of B in their own shared memory; the
// leader issues tcgen05.mma.cta_group::2 (M = 256) and commits with multicast to both CTAs' mbarriers; each CTA reads
// its own 128 accumulator rows back.  Checks numerics (with a shifted A start address) and cycles per MMA.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                 "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                 "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr int QROWS = 544;
struct Params { int N; int shift; int iters; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2(Params p, const __nv_bfloat16* gA, const __nv_bfloat16* gB, float* gD, long long* gcycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                         // [8][QROWS][16 B]  this CTA's 128 (+shift) rows
  uint8_t* sB = smem + 8 * QROWS * 16;        // [8][N/2][16 B]    this CTA's half of B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int NH = p.N / 2;
  // A rows of CTA r are global rows r*272 .. (each CTA gets its own 272-row window of a 544-row matrix)
  for (int i = tid; i < 8 * QROWS * 8; i += 128) {
    int j = i / (QROWS * 8), q = (i / 8) % QROWS, e = i % 8;
    reinterpret_cast<__nv_bfloat16*>(sA)[i] = (q < 272) ? gA[((size_t)j * 544 + rank * 272 + q) * 8 + e] : __float2bfloat16(0.f);
  }
  for (int i = tid; i < 8 * NH * 8; i += 128) {
    int j = i / (NH * 8), n = (i / 8) % NH, e = i % 8;
    reinterpret_cast<__nv_bfloat16*>(sB)[i] = gB[((size_t)j * 256 + rank * NH + n) * 8 + e];
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc2(smem_u32(&tmem_base_s), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc(256, p.N);
  uint32_t parity = 0;
  if (rank == 0 && tid == 0) {
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_desc(smem_u32(sA) + (2 * k) * QROWS * 16 + p.shift * 16, QROWS * 16, 128);
      uint64_t bd = make_desc(smem_u32(sB) + (2 * k) * NH * 16, NH * 16, 128);
      umma2_f16(tmem, ad, bd, idesc, k > 0);
    }
    umma2_commit_mc(smem_u32(&bar), 3);
  }
  mbar_wait(smem_u32(&bar), parity);
  parity ^= 1;
  tc_fence_after();
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    for (int j = 0; j < 32; ++j)
      if (c0 + j < p.N) gD[((size_t)rank * 128 + warp * 32 + (tid & 31)) * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (p.iters > 0) {
    if (rank == 0 && tid == 0) {
      const int shifts[9] = {-9, -8, -7, -1, 0, 1, 7, 8, 9};
      long long t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
        uint32_t acc = tmem + (uint32_t)((it & 1) * 256);
        for (int tap = 0; tap < 9; ++tap) {
          int sh = shifts[tap] + 16;
          for (int k = 0; k < 4; ++k) {
            uint64_t ad = make_desc(smem_u32(sA) + (2 * k) * QROWS * 16 + sh * 16, QROWS * 16, 128);
            uint64_t bd = make_desc(smem_u32(sB) + (2 * k) * NH * 16, NH * 16, 128);
            umma2_f16(acc, ad, bd, idesc, (tap | k) > 0);
          }
        }
      }
      long long t1 = clock64();
      umma2_commit_mc(smem_u32(&bar), 3);
      mbar_wait(smem_u32(&bar), parity);
      long long t2 = clock64();
      if (blockIdx.x == 0) { gcycles[0] = t1 - t0; gcycles[1] = t2 - t0; }
    } else if (tid == 0) {
      mbar_wait(smem_u32(&bar), parity);
    }
  }
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tmem, 512);
}

int main(int argc, char** argv) {
  int grid = argc > 1 ? atoi(argv[1]) : 2;
  std::vector<__nv_bfloat16> hA(8 * 544 * 8), hB(8 * 256 * 8);
  auto Aval = [](int q, int k) { return (float)(((q * 3 + k * 5) % 7) - 3); };
  auto Bval = [](int n, int k) { return (float)(((n * 2 + k) % 5) - 2); };
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 544; ++q) for (int e = 0; e < 8; ++e) hA[(j * 544 + q) * 8 + e] = __float2bfloat16(Aval(q, j * 8 + e));
  for (int j = 0; j < 8; ++j) for (int n = 0; n < 256; ++n) for (int e = 0; e < 8; ++e) hB[(j * 256 + n) * 8 + e] = __float2bfloat16(Bval(n, j * 8 + e));
  __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 256 * 256 * 4)); CK(cudaMalloc(&dC, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const int smem_bytes = 8 * QROWS * 16 + 8 * 128 * 16;
  CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  std::vector<float> hD(256 * 256);
  for (int N : {64, 48, 128, 256}) {
    for (int shift : {0, 9}) {
      Params p{N, shift, 64};
      CK(cudaMemset(dD, 0xFF, 256 * 256 * 4)); CK(cudaMemset(dC, 0, 64));
      probe2<<<grid, 128, smem_bytes>>>(p, dA, dB, dD, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d shift=%d: KERNEL FAILED %s\n", N, shift, cudaGetErrorString(e)); return 2; }
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      long long cyc[2]; CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
      int bad = 0; double maxerr = 0;
      for (int r = 0; r < 2; ++r) for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        float ref = 0; for (int k = 0; k < 64; ++k) ref += Aval(r * 272 + m + shift, k) * Bval(n, k);
        double err = fabs((double)ref - (double)hD[(r * 128 + m) * 256 + n]);
        if (!(err <= 1e-3)) ++bad; if (err > maxerr || err != err) maxerr = err;
      }
      double mmas = 64.0 * 36;
      printf("2CTA M=256 N=%3d shift=%d grid=%d : mismatches %5d / %5d maxerr %.3g | issue %.1f cyc/MMA complete %.1f cyc/MMA\n", N, shift, grid, bad, 256 * N, maxerr, cyc[0] / mmas, cyc[1] / mmas);
      if (bad) { printf("  D[0][0..3]= %g %g %g %g ; D[128][0..3]= %g %g %g %g\n", hD[0], hD[1], hD[2], hD[3], hD[128*256], hD[128*256+1], hD[128*256+2], hD[128*256+3]); }
    }
  }
  printf("PROBE2 DONE\n");
  return 0;
}
