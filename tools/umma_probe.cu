// umma_probe.cu — hardware probe for the tcgen05 evaluator design (not part of the product).
// Checks, on a real B200:
//   T1  tcgen05.mma kind::f16, M=128, K-major NO-SWIZZLE smem descriptors, with the A operand's start
//       address shifted by an arbitrary number of rows (the implicit-GEMM 3x3 tap trick);
//   T2  the same for N = 48 / 32 / 16;
//   T3  issue-to-completion cycles per MMA (SS mode) for several N;
//   T4  latency of an 8 KB cp.async.bulk global->shared copy;
//   T5  tcgen05.ld 32x32b lane/column mapping.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#ifndef GRID_DEF
#define GRID_DEF 1
#endif
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
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

// K-major, no swizzle.  addr/LBO/SBO in bytes.
__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f32 accum, bf16 x bf16, K-major both
}

#ifndef QROWS_DEF
#define QROWS_DEF 160
#endif
constexpr int QROWS = QROWS_DEF;   // rows of the A buffer
#ifndef NMAX_DEF
#define NMAX_DEF 256
#endif
constexpr int NMAX = NMAX_DEF;

struct Params {
  int N;          // MMA N
  int shift;      // A row shift
  int swap_lbo;   // 1: swap the LBO/SBO roles
  int iters;      // timing iterations (0 = correctness only)
  int shifted_timing;  // use 9 different tap shifts in the timing loop
};

__global__ void __launch_bounds__(128) probe(Params p, const __nv_bfloat16* gA, const __nv_bfloat16* gB, float* gD, long long* gcycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                // [8][QROWS][16 B]
  uint8_t* sB = smem + 8 * QROWS * 16;               // [8][NMAX][16 B]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill smem (generic proxy), then make it visible to the async proxy
  for (int i = tid; i < 8 * QROWS * 8; i += 128) reinterpret_cast<__nv_bfloat16*>(sA)[i] = gA[i];
  for (int i = tid; i < 8 * NMAX * 8; i += 128) reinterpret_cast<__nv_bfloat16*>(sB)[i] = gB[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc(128, p.N);
  const uint32_t a_lbo = p.swap_lbo ? 128 : QROWS * 16, a_sbo = p.swap_lbo ? QROWS * 16 : 128;
  const uint32_t b_lbo = p.swap_lbo ? 128 : NMAX * 16, b_sbo = p.swap_lbo ? NMAX * 16 : 128;
  uint32_t parity = 0;
  if (tid == 0) {
    // correctness: D = A[shift.., 0:64] * B[0:N, 0:64]^T  -> 4 MMAs of K=16 (2 chunks each)
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_desc(smem_u32(sA) + (2 * k) * QROWS * 16 + p.shift * 16, a_lbo, a_sbo);
      uint64_t bd = make_desc(smem_u32(sB) + (2 * k) * NMAX * 16, b_lbo, b_sbo);
      umma_f16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), parity);
  parity ^= 1;
  tc_fence_after();
  // read back: warp w owns TMEM lanes 32w..32w+31
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    for (int j = 0; j < 32; ++j)
      if (c0 + j < p.N && blockIdx.x == 0) gD[(warp * 32 + (tid & 31)) * NMAX + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.iters > 0 && tid == 0) {
    const int shifts[9] = {-9, -8, -7, -1, 0, 1, 7, 8, 9};
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      uint32_t acc = tmem + (uint32_t)((it & 1) * 256);
      for (int tap = 0; tap < 9; ++tap) {
        int sh = p.shifted_timing ? shifts[tap] + 16 : 16;
        for (int k = 0; k < 4; ++k) {
          uint64_t ad = make_desc(smem_u32(sA) + (2 * k) * QROWS * 16 + sh * 16, a_lbo, a_sbo);
          uint64_t bd = make_desc(smem_u32(sB) + (2 * k) * NMAX * 16, b_lbo, b_sbo);
          umma_f16(acc, ad, bd, idesc, (tap | k) > 0);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), parity);
    long long t2 = clock64();
    if (blockIdx.x == 0) { gcycles[0] = t1 - t0; gcycles[1] = t2 - t0; }
  }
  // T5: how far can the issuing thread run ahead of the tensor pipe?  Time stamp after each of 40 back-to-back issues.
  if (p.iters < 0 && tid == 0) {
    uint64_t ad = make_desc(smem_u32(sA) + 16 * 16, a_lbo, a_sbo);
    uint64_t bd = make_desc(smem_u32(sB), b_lbo, b_sbo);
    long long ts[41];
    ts[0] = clock64();
#pragma unroll
    for (int i = 0; i < 40; ++i) {
      umma_f16(tmem, ad, bd, idesc, 1);
      ts[i + 1] = clock64();
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), parity);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      for (int i = 0; i <= 40; ++i) gcycles[i] = ts[i] - ts[0];
      gcycles[41] = t2 - ts[0];
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// T4: bulk copy latency
__global__ void bulk_probe(const uint8_t* src, int bytes, int outstanding, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int rep = 0; rep < 6; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < outstanding; ++i) {
        mbar_expect_tx(smem_u32(&bars[i]), bytes);
        bulk_g2s(smem_u32(smem) + i * bytes, src + (size_t)((rep * 8 + i) % 16) * bytes, bytes, smem_u32(&bars[i]));
      }
      long long t1 = clock64();
      for (int i = 0; i < outstanding; ++i) mbar_wait(smem_u32(&bars[i]), rep & 1);
      long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  std::vector<__nv_bfloat16> hA(8 * QROWS * 8), hB(8 * NMAX * 8);
  auto Aval = [](int q, int k) { return (float)(((q * 3 + k * 5) % 7) - 3); };
  auto Bval = [](int n, int k) { return (float)(((n * 2 + k) % 5) - 2); };
  for (int j = 0; j < 8; ++j)
    for (int q = 0; q < QROWS; ++q)
      for (int e = 0; e < 8; ++e) hA[(j * QROWS + q) * 8 + e] = __float2bfloat16(Aval(q, j * 8 + e));
  for (int j = 0; j < 8; ++j)
    for (int n = 0; n < NMAX; ++n)
      for (int e = 0; e < 8; ++e) hB[(j * NMAX + n) * 8 + e] = __float2bfloat16(Bval(n, j * 8 + e));
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, 128 * NMAX * 4));
  CK(cudaMalloc(&dC, 64 * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const int smem_bytes = 8 * QROWS * 16 + 8 * NMAX * 16;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  std::vector<float> hD(128 * NMAX);
  auto run = [&](Params p, const char* tag) {
    CK(cudaMemset(dD, 0xFF, 128 * NMAX * 4));
    CK(cudaMemset(dC, 0, 64 * 8));
    probe<<<GRID_DEF, 128, smem_bytes>>>(p, dA, dB, dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL FAILED %s\n", tag, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    long long cyc[2];
    CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < p.N; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += Aval(m + p.shift, k) * Bval(n, k);
        double err = fabs((double)ref - (double)hD[m * NMAX + n]);
        if (!(err <= 1e-3)) ++bad;
        if (err > maxerr || err != err) maxerr = err;
      }
    printf("%-28s N=%3d shift=%2d swap=%d : mismatches %5d / %5d  maxerr %.3g", tag, p.N, p.shift, p.swap_lbo, bad, 128 * p.N, maxerr);
    if (p.iters) {
      double mmas = (double)p.iters * 36;
      printf("  | timing: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%d MMAs)", cyc[0] / mmas, cyc[1] / mmas, (int)mmas);
    }
    printf("\n");
    if (bad && p.N == 64 && p.shift == 0) {
      printf("   sample D[0][0..7]:");
      for (int n = 0; n < 8; ++n) printf(" %g", hD[n]);
      printf("\n   expect D[0][0..7]:");
      for (int n = 0; n < 8; ++n) { float ref = 0; for (int k = 0; k < 64; ++k) ref += Aval(0, k) * Bval(n, k); printf(" %g", ref); }
      printf("\n");
    }
    return bad;
  };
  int swap = 0;   // confirmed on hardware: LBO = stride between K chunks, SBO = stride between 8-row groups
  run({64, 0, swap, 0, 0}, "T1 base");
  for (int sh : {1, 7, 8, 9, 17, 25}) run({64, sh, swap, 0, 0}, "T1 shifted A");
  for (int N : {16, 32, 48, 128, 256}) if (N <= NMAX) run({N, 3, swap, 0, 0}, "T2 other N");
  for (int N : {16, 32, 48, 64, 128, 256}) { if (N > NMAX) continue;
    run({N, 0, swap, 64, 0}, "T3 timing same-A");
    run({N, 0, swap, 64, 1}, "T3 timing shifted taps");
  }
  // T5
  for (int N : {64, 256}) { if (N > NMAX) continue;
    Params p{N, 0, swap, -1, 0};
    CK(cudaMemset(dC, 0, 64 * 8));
    probe<<<1, 128, smem_bytes>>>(p, dA, dB, dD, dC);
    CK(cudaDeviceSynchronize());
    long long c[42];
    CK(cudaMemcpy(c, dC, sizeof c, cudaMemcpyDeviceToHost));
    printf("T5 issue-queue depth N=%3d: cycles at which MMA i had been issued (i=1..40):", N);
    for (int i = 1; i <= 40; ++i) printf(" %lld", c[i]);
    printf(" | all complete at %lld\n", c[41]);
  }
  // T4
  uint8_t* dsrc;
  CK(cudaMalloc(&dsrc, 16 * 32768));
  CK(cudaMemset(dsrc, 1, 16 * 32768));
  CK(cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
  for (int bytes : {2048, 8192, 16384})
    for (int outst : {1, 4, 8}) {
      CK(cudaMemset(dC, 0, 64 * 8));
      bulk_probe<<<1, 32, 8 * 16384>>>(dsrc, bytes, outst, dC);
      CK(cudaDeviceSynchronize());
      long long c[12];
      CK(cudaMemcpy(c, dC, sizeof c, cudaMemcpyDeviceToHost));
      printf("T4 bulk copy %5d B x %d outstanding: issue/complete cycles per rep:", bytes, outst);
      for (int r = 0; r < 6; ++r) printf(" %lld/%lld", c[2 * r], c[2 * r + 1]);
      printf("\n");
    }
  printf("PROBE DONE\n");
  return 0;
}
