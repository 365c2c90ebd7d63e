// evaluator_umma_v1.cu — the fused policy/value network on tcgen05 tensor cores (sm_100a): one MMA group per 3x3 tap (N = 64).
// This is the FIRST tcgen05 evaluator, kept as a cross-check (SPB_FLAG_EVAL_V1).  The default is evaluator_umma_v2.cu.
//
// Replaces Net::forward + softmax of the reference (ref: src/model/connect_four.rs:75-81,
// src/model/tictactoe.rs:75-81, src/model/mod.rs:62-63; layers model/mod.rs:152-184,
// connect_four.rs:50-73) for a batch of gathered leaf positions.
//
// One persistent CTA per SM.  A CTA takes up to NB boards at a time and runs the WHOLE network on
// them with the activations resident in shared memory:
//
//   positions  a board is laid out as RP rows of W8 cells (Connect4: 7 x 8, one zero pad column and one
//              zero pad row), boards back to back: row index m = board*BS + r*W8 + c.  With this layout a
//              3x3 tap (dy,dx) is a constant row offset dy*W8+dx, and the zero pad cells give the conv's
//              zero padding for free.  NB boards = 4 tiles of 128 rows.
//   A operand  activations, bf16, K-major, NO swizzle: [8 channel-chunks][Q rows][8 channels] so that a core
//              matrix (8 rows x 16 B) is contiguous at ANY row offset -> the tap shift is just a different
//              start address in the shared-memory descriptor (checked on hardware: tools/umma_probe.cu).
//   B operand  weights of one tap, bf16 [8 chunks][N out-channels][8], streamed per layer into a 9-slot ring
//              by cp.async.bulk (TMA, 1-D) from an L2-resident image; BatchNorm is folded in on the host.
//   D          fp32 accumulators in TMEM, one 64-column block per tile (4 x 64 = 256 columns).
//   MMA        tcgen05.mma.cta_group::1.kind::f16, M=128, N=64 (48 for the fused policy+value head conv),
//              K=16; 9 taps x 4 k-steps accumulate one tile of one layer.
//   epilogue   8 warps (two per TMEM lane quadrant, 32 channels each): tcgen05.ld -> +bias (+skip) -> ReLU -> bf16 -> st.shared into the other activation
//              buffer (pad rows forced to zero), which is the next layer's A operand.  Head layer: the two
//              Linear layers, softmax and tanh are computed from the accumulators.
//
// Warp roles: warp 0 = weight producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..9 = epilogue (TMEM lane quadrant = warp & 3, channel half = (warp-2)/4).  Hand-offs are mbarriers only.
#include <cuda_bf16.h>

#include <cstring>
#include <mutex>

#include "evaluator_umma.cuh"

namespace spb {
namespace umma_v1 {

// ---------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------
template <class G>
struct Geo {
  static constexpr int W8 = (G::COLS == 7) ? 8 : 4;          // padded row width
  static constexpr int RP = G::ROWS + 1;                     // rows incl. the shared zero pad row
  static constexpr int BS = W8 * RP;                         // rows per board (56 / 16)
  static constexpr int NT = 4;                               // tiles per batch
  static constexpr int NB = (NT * 128) / BS;                 // boards per batch (9 / 32)
  static constexpr int LEAD = 16;                            // zero rows in front (taps reach back W8+1)
  static constexpr int Q = LEAD + NT * 128 + 16;             // rows of an activation buffer
  static constexpr int P = G::ROWS * G::COLS;
  static constexpr int APAD = (G::A <= 8) ? 8 : 16;          // policy FC weights per (pos, channel), bf16
};

constexpr int N_LAYERS = 10;          // stem, 8 residual convs, fused head conv
constexpr int HEAD_N = 48;            // 32 policy + 3 value + 13 zero output channels
constexpr int SLOT_BYTES = 8192;      // one tap of a 64->64 layer
constexpr int N_SLOTS = 9;

__host__ __device__ constexpr int layer_n(int l) { return l == 9 ? HEAD_N : 64; }
__host__ __device__ constexpr int layer_kchunks(int l) { return l == 0 ? 2 : 8; }
__host__ __device__ constexpr int layer_tap_bytes(int l) { return layer_kchunks(l) * layer_n(l) * 16; }
__host__ __device__ constexpr size_t layer_offset(int l) {
  return l == 0 ? 0 : (size_t)9 * 2048 + (size_t)(l - 1) * 9 * SLOT_BYTES;
}
constexpr size_t OFF_BIAS = (size_t)9 * 2048 + (size_t)8 * 9 * SLOT_BYTES + (size_t)9 * 6144;   // 663,552
constexpr size_t OFF_WP = OFF_BIAS + (size_t)N_LAYERS * 64 * 4;
template <class G> __host__ __device__ constexpr size_t off_wv() { return OFF_WP + (size_t)Geo<G>::P * 32 * Geo<G>::APAD * 2; }
template <class G> __host__ __device__ constexpr size_t off_fcb() { return off_wv<G>() + (size_t)Geo<G>::P * 4 * 4; }
template <class G> __host__ __device__ constexpr size_t image_bytes() { return off_fcb<G>() + 32 * 4; }

// ---------------------------------------------------------------------------------------------------
// host: weight image
// ---------------------------------------------------------------------------------------------------
static inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

template <class G>
static void pack_t(const HostNet& net, std::vector<uint8_t>* out) {
  using Ge = Geo<G>;
  out->assign(image_bytes<G>(), 0);
  uint8_t* img = out->data();
  for (int l = 0; l < N_LAYERS; ++l) {
    const int N = layer_n(l), KC = layer_kchunks(l);
    for (int tap = 0; tap < 9; ++tap) {
      uint16_t* blk = reinterpret_cast<uint16_t*>(img + layer_offset(l) + (size_t)tap * layer_tap_bytes(l));
      for (int n = 0; n < N; ++n) {
        const HostNet::Conv* cv;
        int oc;
        if (l < 9) { cv = &net.conv[l]; oc = n; }
        else if (n < NET_POLICY_CH) { cv = &net.conv[9]; oc = n; }
        else if (n < NET_POLICY_CH + NET_VALUE_CH) { cv = &net.conv[10]; oc = n - NET_POLICY_CH; }
        else continue;
        for (int k = 0; k < KC * 8; ++k) {
          if (k >= cv->ic) break;
          float w = cv->w[((size_t)oc * cv->ic + k) * 9 + tap];   // [OC][IC][ky][kx], tap = ky*3+kx
          blk[((size_t)(k / 8) * N + n) * 8 + (k % 8)] = f2bf(w);
        }
      }
    }
  }
  float* bias = reinterpret_cast<float*>(img + OFF_BIAS);
  for (int l = 0; l < 9; ++l)
    for (int n = 0; n < 64; ++n) bias[l * 64 + n] = net.conv[l].b[n];
  for (int n = 0; n < NET_POLICY_CH; ++n) bias[9 * 64 + n] = net.conv[9].b[n];
  for (int n = 0; n < NET_VALUE_CH; ++n) bias[9 * 64 + NET_POLICY_CH + n] = net.conv[10].b[n];
  // policy Linear: weight[a][ch*P + pos] -> bf16 [pos][ch][APAD]
  uint16_t* wp = reinterpret_cast<uint16_t*>(img + OFF_WP);
  for (int pos = 0; pos < Ge::P; ++pos)
    for (int ch = 0; ch < NET_POLICY_CH; ++ch)
      for (int a = 0; a < G::A; ++a)
        wp[((size_t)pos * NET_POLICY_CH + ch) * Ge::APAD + a] = f2bf(net.pfc_w[(size_t)a * NET_POLICY_CH * Ge::P + (size_t)ch * Ge::P + pos]);
  // value Linear: weight[0][ch*P + pos] -> f32 [pos][4]
  float* wv = reinterpret_cast<float*>(img + off_wv<G>());
  for (int pos = 0; pos < Ge::P; ++pos)
    for (int ch = 0; ch < NET_VALUE_CH; ++ch) wv[pos * 4 + ch] = net.vfc_w[(size_t)ch * Ge::P + pos];
  float* fcb = reinterpret_cast<float*>(img + off_fcb<G>());
  for (int a = 0; a < G::A; ++a) fcb[a] = net.pfc_b[a];
  fcb[16] = net.vfc_b[0];
}

void pack_weights(const HostNet& net, std::vector<uint8_t>* out) {
  if (net.game == SPB_GAME_CONNECT4) pack_t<Connect4>(net, out);
  else pack_t<TicTacToe>(net, out);
}

// ---------------------------------------------------------------------------------------------------
// device helpers (inline PTX)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {   // for the producer: don't hog issue slots
  uint32_t ok;
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(200);
  }
}
template <int NS>
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {   // waiters with slack: poll less, leave the
  uint32_t ok;                                                                         // shared-memory pipe to the tensor core
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (NS > 0) __nanosleep(NS);
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// One elected lane of a fully converged warp (cute::elect_one_sync): lets the compiler keep tcgen05 operands in
// uniform registers instead of emitting a per-lane waterfall loop around every instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor: LBO = byte stride between the two 8-element K
// chunks of one MMA, SBO = byte stride between 8-row groups (verified by tools/umma_probe.cu).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Issues the 9 taps x KSTEPS MMAs of one (layer, tile).  Descriptor low words: A = a_lo_tile + tap shift +
// kk * 2Q (two K chunks further), B = slot base + tap * slot stride + kk * 2N; high words are constants.
template <int W8, int Q, int KSTEPS, int N>
__device__ __forceinline__ void issue_tile(bool issuer, uint32_t a_lo_tile, uint32_t b_lo_base, uint32_t slot_stride16,
                                           uint32_t d_tmem, bool first_tile, bool last_tile, uint32_t w_par, uint32_t bar_base,
                                           unsigned long long& prof_wfull) {
  constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);          // SBO = 128 B, descriptor version 1
  constexpr uint32_t IDESC = make_idesc(N);
  (void)slot_stride16;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    if (first_tile) {                                                                            // w_full[tap]
#ifdef SPB_PROFILE
      const unsigned long long t0 = clock64();
#endif
      mbar_wait(bar_base + (uint32_t)tap * 8u, w_par);
#ifdef SPB_PROFILE
      prof_wfull += clock64() - t0;
#endif
      tc_fence_after();
    }
    const int shift = (tap / 3 - 1) * W8 + (tap % 3 - 1);
    if (issuer) {
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const uint32_t a_lo = a_lo_tile + (uint32_t)(shift + kk * 2 * Q);
        const uint32_t b_lo = (b_lo_base + (uint32_t)tap * (SLOT_BYTES >> 4) + (uint32_t)(kk * 2 * N)) | ((uint32_t)N << 16);
        const uint64_t ad = ((uint64_t)DESC_HI << 32) | a_lo;
        const uint64_t bd = ((uint64_t)DESC_HI << 32) | b_lo;
        umma_f16(d_tmem, ad, bd, IDESC, (tap | kk) != 0);
      }
      if (last_tile) umma_commit(bar_base + (uint32_t)(N_SLOTS + tap) * 8u);                    // w_empty[tap]
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// shared memory plan
// ---------------------------------------------------------------------------------------------------
template <class G>
struct Smem {
  using Ge = Geo<G>;
  static constexpr int ACT_BYTES = 8 * Ge::Q * 16;                 // 69,632
  static constexpr int OFF_ACT0 = 0;
  static constexpr int OFF_ACT1 = ACT_BYTES;
  static constexpr int OFF_W = 2 * ACT_BYTES;                      // 9 x 8 KB weight ring
  static constexpr int OFF_BIAS = OFF_W + N_SLOTS * SLOT_BYTES;    // 10 x 64 f32
  static constexpr int OFF_LOGITS = OFF_BIAS + N_LAYERS * 64 * 4;  // [NB][16] f32 (policy partial sums, value at [15])
  static constexpr int OFF_PART = OFF_LOGITS + Ge::NB * 16 * 4;    // [NB][ROWS][16] f32 row partials of the Linear layers
  static constexpr int OFF_STATES = OFF_PART + Ge::NB * G::ROWS * 16 * 4;   // [2][NB] PState
  static constexpr int OFF_SLOTS = OFF_STATES + 2 * Ge::NB * 16;   // [2][NB] u32 (states/slots ping-pong per batch)
  static constexpr int OFF_BARS = (OFF_SLOTS + 2 * Ge::NB * 4 + 15) & ~15;
  // barriers: w_full[9], w_empty[9], acc_full[4], act_ready[4], stage_ready[4], acc_full of odd batches [4]
  static constexpr int OFF_TMEM = OFF_BARS + (2 * N_SLOTS + 4 * Ge::NT) * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
};

static_assert(Smem<Connect4>::TOTAL <= 232448 && Smem<TicTacToe>::TOTAL <= 232448, "shared memory plan exceeds 227 KB");
constexpr int THREADS = 320;     // producer warp, MMA warp, 8 epilogue warps

#ifdef SPB_PROFILE
// debug build only (make PROFILE=1): per-CTA cycle attribution
__device__ unsigned long long g_eval_prof[160][8];
__device__ unsigned long long g_eval_prof_layer[160][24];   // [cta][0..9] act waits per layer, [10..19] weight waits per layer
__device__ int g_eval_debug = 0;   // bit0: epilogue skips tcgen05.ld, bit1: skips st.shared, bit2: skips skip-loads, bit3: no per-tap commits
#define DBG(bit) (g_eval_debug & (1 << (bit)))
#define PROF_DECL unsigned long long prof_t0 = 0, prof_acc0 = 0, prof_acc1 = 0, prof_acc2 = 0;
#define PROF_BEGIN() (prof_t0 = clock64())
#define PROF_END(acc) ((acc) += clock64() - prof_t0)
#else
#define DBG(bit) 0
#define PROF_DECL
#define PROF_BEGIN() ((void)0)
#define PROF_END(acc) ((void)0)
#endif

template <class G>
__global__ void __launch_bounds__(THREADS, 1)
k_eval_umma(const uint8_t* __restrict__ image, const PState* __restrict__ states, const uint32_t* __restrict__ list,
            const uint32_t* __restrict__ count_dev, uint32_t max_n, float* __restrict__ out, int stride,
            float* __restrict__ logits_out) {
  using Ge = Geo<G>;
  using Sm = Smem<G>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t n_total = min(*count_dev, max_n);
  const uint32_t cta = blockIdx.x, ncta = gridDim.x;
  const uint32_t my_begin = (uint32_t)(((uint64_t)n_total * cta) / ncta);
  const uint32_t my_end = (uint32_t)(((uint64_t)n_total * (cta + 1)) / ncta);
  if (my_begin >= my_end) return;                                  // uniform per CTA

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar_base = s_base + Sm::OFF_BARS;
  auto bar_w_full = [&](int s) { return bar_base + (uint32_t)s * 8u; };
  auto bar_w_empty = [&](int s) { return bar_base + (uint32_t)(N_SLOTS + s) * 8u; };
  // acc_full is per accumulator set (batch parity): the MMA warp may finish the next batch's stem tile before the
  // epilogue has consumed this batch's head tile, and an mbarrier must never run two phases ahead of a waiter.
  auto bar_acc_full = [&](uint32_t set, int t) { return bar_base + (uint32_t)(2 * N_SLOTS + (set ? 3 * Ge::NT : 0) + t) * 8u; };
  auto bar_act_ready = [&](int t) { return bar_base + (uint32_t)(2 * N_SLOTS + Ge::NT + t) * 8u; };
  // separate barrier for the staged input of a batch: it may complete while act_ready's previous phase is still
  // being consumed by the MMA warp (an mbarrier must never run two phases ahead of a waiter)
  auto bar_stage_ready = [&](int t) { return bar_base + (uint32_t)(2 * N_SLOTS + 2 * Ge::NT + t) * 8u; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Sm::OFF_TMEM);

  // ---- one-time setup -----------------------------------------------------------------------------
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* p = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < 2 * Sm::ACT_BYTES / 16; i += THREADS) p[i] = z;           // pad rows stay zero forever
    const float* gb = reinterpret_cast<const float*>(image + OFF_BIAS);
    float* sb = reinterpret_cast<float*>(smem + Sm::OFF_BIAS);
    for (int i = tid; i < N_LAYERS * 64; i += THREADS) sb[i] = gb[i];
  }
  if (tid == 0) {
    for (int s = 0; s < N_SLOTS; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 1); }
    for (int t = 0; t < Ge::NT; ++t) { mbar_init(bar_acc_full(0, t), 1); mbar_init(bar_acc_full(1, t), 1); mbar_init(bar_act_ready(t), 256); mbar_init(bar_stage_ready(t), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t n_batches = (my_end - my_begin + Ge::NB - 1) / Ge::NB;

  if (warp == 0) {
    // ===== weight producer: streams (layer, tap) blocks into the 9-slot ring ==========================
    if (lane == 0) {
      uint32_t use = 0;                                           // completed fills of every slot
      for (uint32_t b = 0; b < n_batches; ++b) {
        for (int l = 0; l < N_LAYERS; ++l) {
          const uint32_t bytes = (uint32_t)layer_tap_bytes(l);
          const uint8_t* src = image + layer_offset(l);
          for (int s = 0; s < N_SLOTS; ++s) {
#ifdef V_PROD_SLEEP
            if (use > 0) mbar_wait_backoff<V_PROD_SLEEP>(bar_w_empty(s), (use - 1) & 1u);
#else
            if (use > 0) mbar_wait(bar_w_empty(s), (use - 1) & 1u);
#endif
            if (DBG(4) && use > 0) { mbar_arrive(bar_w_full(s)); continue; }   // profile build: no weight traffic
            mbar_expect_tx(bar_w_full(s), bytes);
            bulk_g2s(s_base + Sm::OFF_W + (uint32_t)s * SLOT_BYTES, src + (size_t)s * bytes, bytes, bar_w_full(s));
          }
          ++use;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===============================================================================
    // The whole warp runs the (warp-uniform) control flow so that descriptors stay in uniform registers;
    // one fixed lane issues the tcgen05 instructions.
    {
      PROF_DECL
#ifdef SPB_PROFILE
      const unsigned long long prof_start = clock64();
#endif
      unsigned long long prof_w = 0;
      const bool issuer = elect_one();
      const uint32_t b_lo_base = ((s_base + Sm::OFF_W) >> 4);
      uint32_t use = 0;        // layer-uses of the weight ring so far
      uint32_t act_par = 0;    // bit t: parity of the next completion of act_ready[t]
      uint32_t stage_par = 0;  // same for stage_ready[t]
      for (uint32_t b = 0; b < n_batches; ++b) {
        const uint32_t nb = min((uint32_t)Ge::NB, my_end - my_begin - b * Ge::NB);
        const int nt = (int)((nb * Ge::BS + 127) / 128);
        for (int l = 0; l < N_LAYERS; ++l) {
          const uint32_t in_buf = s_base + ((l == 0 || (l >= 2 && (l & 1) == 0)) ? Sm::OFF_ACT0 : Sm::OFF_ACT1);
          const uint32_t a_lo_base = ((in_buf >> 4) + Ge::LEAD) | ((uint32_t)Ge::Q << 16);
          const uint32_t cur_par = (l == 0) ? stage_par : act_par;
          if (l == 0) stage_par ^= (1u << nt) - 1u; else act_par ^= (1u << nt) - 1u;
          for (int t = 0; t < nt; ++t) {
            const int wt = min(t + 1, nt - 1);                   // epilogue runs tiles in order: tile wt done => 0..wt done
            PROF_BEGIN();
            mbar_wait(l == 0 ? bar_stage_ready(wt) : bar_act_ready(wt), (cur_par >> wt) & 1u);
            PROF_END(prof_acc0);
#ifdef SPB_PROFILE
            if (lane == 0) g_eval_prof_layer[blockIdx.x][l] += clock64() - prof_t0;
            const unsigned long long w_before = prof_w;
            prof_t0 = clock64();
#endif
            tc_fence_after();
            const uint32_t a_lo_tile = a_lo_base + (uint32_t)t * 128u;
            const uint32_t d_tmem = tmem_base + (b & 1u) * 256u + (uint32_t)t * 64u;   // accumulators ping-pong per batch
            const bool first = (t == 0), last = (t == nt - 1);
            if (l == 0)
              issue_tile<Ge::W8, Ge::Q, 1, 64>(issuer, a_lo_tile, b_lo_base, 2048 >> 4, d_tmem, first, last, use & 1u, bar_base, prof_w);
            else if (l < 9)
              issue_tile<Ge::W8, Ge::Q, 4, 64>(issuer, a_lo_tile, b_lo_base, SLOT_BYTES >> 4, d_tmem, first, last, use & 1u, bar_base, prof_w);
            else
              issue_tile<Ge::W8, Ge::Q, 4, HEAD_N>(issuer, a_lo_tile, b_lo_base, SLOT_BYTES >> 4, d_tmem, first, last, use & 1u, bar_base, prof_w);
#ifdef SPB_PROFILE
            if (lane == 0) {
              g_eval_prof_layer[blockIdx.x][10 + l] += prof_w - w_before;
              if (l >= 1 && l <= 8 && nt == 4) {   // issue time of one residual-layer tile by position, weight waits excluded
                const unsigned long long dt = clock64() - prof_t0 - (prof_w - w_before);
                g_eval_prof_layer[blockIdx.x][20 + (first ? 0 : last ? 2 : 1)] += dt;
                if (first) g_eval_prof_layer[blockIdx.x][23] += 1;
              }
            }
#endif
            if (issuer) umma_commit(bar_acc_full(b & 1u, t));
            __syncwarp();
          }
          ++use;
        }
      }
#ifdef SPB_PROFILE
      if (lane == 0) {
        g_eval_prof[blockIdx.x][0] = clock64() - prof_start;   // MMA warp total
        g_eval_prof[blockIdx.x][1] = prof_acc0;                // waiting for activations (epilogue)
        g_eval_prof[blockIdx.x][2] = n_batches;
        g_eval_prof[blockIdx.x][5] = prof_w;                    // waiting for weights (TMA ring)
      }
#endif
    }
  } else {
    // ===== epilogue warps (8 warps, 256 threads): encode, per-layer epilogues, heads ====================
    // Two warps share a TMEM lane quadrant (a tile row) and split the 64 output channels in halves.
    const int et = tid - 64;                                       // 0..255
    const int quad = warp & 3;                                     // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;                              // channels [32*half, 32*half+32)
    const int row_in_tile = quad * 32 + lane;
    const float* s_bias = reinterpret_cast<const float*>(smem + Sm::OFF_BIAS);
    float* s_logits = reinterpret_cast<float*>(smem + Sm::OFF_LOGITS);
    float* s_part = reinterpret_cast<float*>(smem + Sm::OFF_PART);
    PState* s_states = reinterpret_cast<PState*>(smem + Sm::OFF_STATES);
    uint32_t* s_slots = reinterpret_cast<uint32_t*>(smem + Sm::OFF_SLOTS);
    const uint16_t* g_wp = reinterpret_cast<const uint16_t*>(image + OFF_WP);
    const float* g_wv = reinterpret_cast<const float*>(image + off_wv<G>());
    const float* g_fcb = reinterpret_cast<const float*>(image + off_fcb<G>());
    constexpr int PCH = (G::A + 1 + 3) / 4;                        // 16-B chunks of head partial sums per half
    uint32_t acc_par[2] = {0, 0};                                  // [set] bit t: parity of the next completion of acc_full[set][t]
    PROF_DECL
#ifdef SPB_PROFILE
    const unsigned long long prof_start = clock64();
#endif

    // Fetches the states of batch `bb` and writes their encoding (get_encoding, connect_four.rs:242-259: channels
    // 0,1,2 of chunk 0; chunk 1 = 0) into activation buffer 0, then releases the stem MMAs.  Called for batch b+1
    // while the tensor pipe runs the head conv of batch b, so the pipe never waits for a batch turn-around.
    auto stage_batch = [&](uint32_t bb) {
      const uint32_t b0 = my_begin + bb * Ge::NB;
      const uint32_t nb = min((uint32_t)Ge::NB, my_end - b0);
      const int nt = (int)((nb * Ge::BS + 127) / 128);
      PState* st_buf = s_states + (bb & 1u) * Ge::NB;
      uint32_t* sl_buf = s_slots + (bb & 1u) * Ge::NB;
      if ((uint32_t)et < nb) {
        const uint32_t slot = list ? list[b0 + et] : (b0 + et);
        sl_buf[et] = slot;
        st_buf[et] = states[slot];
      }
      epi_bar_sync();
      for (int t = 0; t < nt; ++t) {
        const int m = t * 128 + row_in_tile;
        const int bi = m / Ge::BS, rem = m % Ge::BS, r = rem / Ge::W8, c = rem % Ge::W8;
        uint4 v0 = make_uint4(0, 0, 0, 0);
        if (half == 0 && (uint32_t)bi < nb && r < G::ROWS && c < G::COLS) {
          const PState st = st_buf[bi];
          const float e0 = G::encode_cell(st, 0, r, c), e1 = G::encode_cell(st, 1, r, c), e2 = G::encode_cell(st, 2, r, c);
          v0.x = pack_bf16x2(e0, e1);
          v0.y = pack_bf16x2(e2, 0.0f);
        }
        *reinterpret_cast<uint4*>(smem + Sm::OFF_ACT0 + (size_t)half * Ge::Q * 16 + (size_t)(Ge::LEAD + m) * 16) = v0;
        fence_async_smem();
        mbar_arrive(bar_stage_ready(t));
      }
    };
    stage_batch(0);

    for (uint32_t b = 0; b < n_batches; ++b) {
      const uint32_t b0 = my_begin + b * Ge::NB;
      const uint32_t nb = min((uint32_t)Ge::NB, my_end - b0);
      const int nt = (int)((nb * Ge::BS + 127) / 128);
      const uint32_t tmem_acc = tmem_base + (b & 1u) * 256u;
      const uint32_t* sl_cur = s_slots + (b & 1u) * Ge::NB;
      // ---- layers
      for (int l = 0; l < N_LAYERS; ++l) {
        const bool in0 = (l == 0 || (l >= 2 && (l & 1) == 0));
        uint8_t* dst_buf = smem + (in0 ? Sm::OFF_ACT1 : Sm::OFF_ACT0);
        const bool has_skip = (l >= 2 && (l & 1) == 0 && l <= 8);   // second conv of a residual block
        const uint32_t cur_par = acc_par[b & 1u];
        acc_par[b & 1u] ^= (1u << nt) - 1u;
        if (l < 9) {
          float bias_r[32];                                         // this thread's 32 output channels
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + l * 64 + half * 32 + q * 4);
            bias_r[4 * q] = bv.x; bias_r[4 * q + 1] = bv.y; bias_r[4 * q + 2] = bv.z; bias_r[4 * q + 3] = bv.w;
          }
          for (int t = 0; t < nt; ++t) {
            const int m = t * 128 + row_in_tile;
            const int bi = m / Ge::BS, rem = m % Ge::BS, r = rem / Ge::W8, c = rem % Ge::W8;
            const bool valid = (uint32_t)bi < nb && r < G::ROWS && c < G::COLS;
            uint8_t* drow = dst_buf + (size_t)(half * 4) * Ge::Q * 16 + (size_t)(Ge::LEAD + m) * 16;   // chunk 4*half
            uint4 sk[4];
            if (has_skip) {                                         // (x + f(x)).relu(), model/mod.rs:163
#pragma unroll
              for (int j = 0; j < 4; ++j) sk[j] = DBG(2) ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4*>(drow + (size_t)j * Ge::Q * 16);
            }
            PROF_BEGIN();
#ifdef V_EPI_BACKOFF
            mbar_wait_backoff<V_EPI_BACKOFF>(bar_acc_full(b & 1u, t), (cur_par >> t) & 1u);
#else
            mbar_wait(bar_acc_full(b & 1u, t), (cur_par >> t) & 1u);
#endif
            PROF_END(prof_acc0);
            tc_fence_after();
#ifdef SPB_PROFILE
            const unsigned long long body0 = clock64();
#endif
            const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)t * 64u + (uint32_t)half * 32u;
            uint32_t a[32];
            if (DBG(0)) {
#pragma unroll
              for (int q = 0; q < 32; ++q) a[q] = 0;
            } else {
              tmem_ld16(taddr, a);
              tmem_ld16(taddr + 16, a + 16);
              tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                           // one 8-channel chunk = one 16-B store
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(a[j * 8 + e]) + bias_r[j * 8 + e];
              if (has_skip) {
                v[0] += bf_lo(sk[j].x); v[1] += bf_hi(sk[j].x); v[2] += bf_lo(sk[j].y); v[3] += bf_hi(sk[j].y);
                v[4] += bf_lo(sk[j].z); v[5] += bf_hi(sk[j].z); v[6] += bf_lo(sk[j].w); v[7] += bf_hi(sk[j].w);
              }
              uint4 o = make_uint4(0, 0, 0, 0);
              if (valid) {
                o.x = pack_bf16x2(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f));
                o.y = pack_bf16x2(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
                o.z = pack_bf16x2(fmaxf(v[4], 0.f), fmaxf(v[5], 0.f));
                o.w = pack_bf16x2(fmaxf(v[6], 0.f), fmaxf(v[7], 0.f));
              }
              if (!DBG(1)) *reinterpret_cast<uint4*>(drow + (size_t)j * Ge::Q * 16) = o;
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_act_ready(t));
#ifdef SPB_PROFILE
            prof_acc1 += clock64() - body0;
            prof_acc2 += 1;
#endif
          }
        } else {
          // The head conv's MMAs are in flight: stage the next batch now (buffer 0 is no longer read by this batch).
          if (b + 1 < n_batches) stage_batch(b + 1);
          // ---- heads: policy conv channels 0..31 (16 per half), value conv channels 32..34 (half 1), then the
          // per-position terms of the two Linear layers.
          const float* bias = s_bias + l * 64;
          for (int t = 0; t < nt; ++t) {
            const int m = t * 128 + row_in_tile;
            const int bi = m / Ge::BS, rem = m % Ge::BS, r = rem / Ge::W8, c = rem % Ge::W8;
            const bool valid = (uint32_t)bi < nb && r < G::ROWS && c < G::COLS;
            const int pos = r * G::COLS + c;
            // prefetch this position's policy-Linear weights (bf16 [pos][ch][APAD]) before waiting for the MMA
            uint4 w[16 * (Ge::APAD / 8)];
            if (valid) {
              const uint4* wrow = reinterpret_cast<const uint4*>(g_wp + ((size_t)pos * NET_POLICY_CH + half * 16) * Ge::APAD);
#pragma unroll
              for (int i = 0; i < 16 * (Ge::APAD / 8); ++i) w[i] = __ldg(wrow + i);
            }
            PROF_BEGIN();
            mbar_wait(bar_acc_full(b & 1u, t), (cur_par >> t) & 1u);
            PROF_END(prof_acc0);
            tc_fence_after();
            const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)t * 64u;
            uint32_t a[16], av[16];
            tmem_ld16(taddr + (uint32_t)half * 16u, a);
            if (half == 1) tmem_ld16(taddr + 32u, av);
            tmem_ld_wait();
            tc_fence_before();
            if (valid) {
              float pl[Ge::APAD + 4];
#pragma unroll
              for (int i = 0; i < Ge::APAD + 4; ++i) pl[i] = 0.0f;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float act = fmaxf(__uint_as_float(a[e]) + bias[half * 16 + e], 0.0f);
#pragma unroll
                for (int q = 0; q < Ge::APAD / 8; ++q) {
                  const uint4 wq = w[e * (Ge::APAD / 8) + q];
                  pl[q * 8 + 0] = fmaf(act, bf_lo(wq.x), pl[q * 8 + 0]); pl[q * 8 + 1] = fmaf(act, bf_hi(wq.x), pl[q * 8 + 1]);
                  pl[q * 8 + 2] = fmaf(act, bf_lo(wq.y), pl[q * 8 + 2]); pl[q * 8 + 3] = fmaf(act, bf_hi(wq.y), pl[q * 8 + 3]);
                  pl[q * 8 + 4] = fmaf(act, bf_lo(wq.z), pl[q * 8 + 4]); pl[q * 8 + 5] = fmaf(act, bf_hi(wq.z), pl[q * 8 + 5]);
                  pl[q * 8 + 6] = fmaf(act, bf_lo(wq.w), pl[q * 8 + 6]); pl[q * 8 + 7] = fmaf(act, bf_hi(wq.w), pl[q * 8 + 7]);
                }
              }
              float vl = 0.0f;
              if (half == 1) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(g_wv) + pos);
                vl = fmaf(fmaxf(__uint_as_float(av[0]) + bias[32], 0.0f), wv.x, vl);
                vl = fmaf(fmaxf(__uint_as_float(av[1]) + bias[33], 0.0f), wv.y, vl);
                vl = fmaf(fmaxf(__uint_as_float(av[2]) + bias[34], 0.0f), wv.z, vl);
              }
              pl[G::A] = vl;
              // Per-position partial sums go to this row's own (now dead) cells of activation buffer 0: chunks
              // 2+PCH*half.. — pad rows stay zero, and every chunk >= 2 is rewritten by layer 1 before it is read again.
              uint8_t* prow = smem + Sm::OFF_ACT0 + (size_t)(Ge::LEAD + m) * 16;
#pragma unroll
              for (int q = 0; q < PCH; ++q)
                *reinterpret_cast<float4*>(prow + (size_t)(2 + PCH * half + q) * Ge::Q * 16) = make_float4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
            }
          }
        }
      }
      // ---- finish the batch
      epi_bar_sync();
      // Linear layers: sum the per-position partials of each board in a FIXED order (deterministic results,
      // independent of how leaves were batched), in two levels: (board, row, output) over the columns and both
      // channel halves, then (board, output) over the rows.
      for (int i = et; i < (int)nb * G::ROWS * 16; i += 256) {
        const int a = i & 15, br = i >> 4, bi = br / G::ROWS, r = br % G::ROWS;
        if (a <= G::A) {
          float acc = 0.0f;
#pragma unroll
          for (int c = 0; c < G::COLS; ++c) {
            const uint8_t* prow = smem + Sm::OFF_ACT0 + (size_t)(Ge::LEAD + bi * Ge::BS + r * Ge::W8 + c) * 16 + (a & 3) * 4;
            acc += *reinterpret_cast<const float*>(prow + (size_t)(2 + (a >> 2)) * Ge::Q * 16);
            acc += *reinterpret_cast<const float*>(prow + (size_t)(2 + PCH + (a >> 2)) * Ge::Q * 16);
          }
          s_part[br * 16 + a] = acc;
        }
      }
      epi_bar_sync();
      for (int i = et; i < (int)nb * 16; i += 256) {
        const int bi = i >> 4, a = i & 15;
        if (a <= G::A) {
          float acc = 0.0f;
#pragma unroll
          for (int r = 0; r < G::ROWS; ++r) acc += s_part[(bi * G::ROWS + r) * 16 + a];
          s_logits[bi * 16 + (a == G::A ? 15 : a)] = acc;
        }
      }
      epi_bar_sync();
      // softmax (model/mod.rs:63) and tanh (connect_four.rs:71), one thread per board
      if ((uint32_t)et < nb) {
        const uint32_t slot = sl_cur[et];
        float lg[G::A];
        float mx = -INFINITY;
#pragma unroll
        for (int a = 0; a < G::A; ++a) { lg[a] = s_logits[et * 16 + a] + g_fcb[a]; mx = fmaxf(mx, lg[a]); }
        float ex[G::A], sum = 0.0f;
#pragma unroll
        for (int a = 0; a < G::A; ++a) { ex[a] = expf(lg[a] - mx); sum += ex[a]; }
        float* o = out + (size_t)slot * stride;
#pragma unroll
        for (int a = 0; a < G::A; ++a) o[a] = ex[a] / sum;
        o[G::A] = tanhf(s_logits[et * 16 + 15] + g_fcb[16]);
        if (logits_out) {
#pragma unroll
          for (int a = 0; a < G::A; ++a) logits_out[(size_t)slot * G::A + a] = lg[a];
        }
      }
      epi_bar_sync();                                              // s_logits / s_states are reused by the next batch
    }
#ifdef SPB_PROFILE
    if (et == 0) {
      g_eval_prof[blockIdx.x][3] = clock64() - prof_start;     // epilogue warp total
      g_eval_prof[blockIdx.x][4] = prof_acc0;                  // waiting for accumulators (MMA)
      g_eval_prof[blockIdx.x][6] = prof_acc1;                  // conv-layer epilogue bodies (after the wait, through the arrive)
      g_eval_prof[blockIdx.x][7] = prof_acc2;                  // number of such bodies
    }
#endif
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <class G>
static cudaError_t launch_t(const Evaluator::DevNet& net, const PState* states, const uint32_t* list, const uint32_t* count_dev,
                            uint32_t max_n, float* out, int stride, float* logits_out, cudaStream_t stream) {
  // per device (one process may drive one engine per GPU from several host threads): SM count + opt-in shared memory
  static std::mutex mu;
  static int sm_counts[64] = {};
  int dev = 0, sm_count = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (sm_counts[dev] == 0) {
      int n = 0;
      e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval_umma<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<G>::TOTAL);
      if (e != cudaSuccess) return e;
      sm_counts[dev] = n;
    }
    sm_count = sm_counts[dev];
  }
  const unsigned grid = (unsigned)std::max(1, std::min<int>(sm_count, (int)max_n));
  k_eval_umma<G><<<grid, THREADS, Smem<G>::TOTAL, stream>>>(reinterpret_cast<const uint8_t*>(net.w_umma), states, list, count_dev, max_n,
                                                            out, stride, logits_out);
  return cudaGetLastError();
}

#ifdef SPB_PROFILE
extern "C" int spb_debug_set_v1(int v) { return (int)cudaMemcpyToSymbol(g_eval_debug, &v, sizeof v); }
extern "C" int spb_debug_eval_profile_layers_v1(unsigned long long* out, int n_ctas, int reset) {
  int rc = (int)cudaMemcpyFromSymbol(out, g_eval_prof_layer, sizeof(unsigned long long) * 24 * (size_t)n_ctas);
  if (reset) { static unsigned long long z[160 * 24]; rc |= (int)cudaMemcpyToSymbol(g_eval_prof_layer, z, sizeof z); }
  return rc;
}
extern "C" int spb_debug_eval_profile_v1(unsigned long long* out, int n_ctas) {
  return (int)cudaMemcpyFromSymbol(out, g_eval_prof, sizeof(unsigned long long) * 8 * (size_t)n_ctas);
}
#endif

cudaError_t launch(const Evaluator::DevNet& net, int game, const PState* states, const uint32_t* list, const uint32_t* count_dev,
                   uint32_t max_n, float* out, int stride, float* logits_out, cudaStream_t stream) {
  if (game == SPB_GAME_CONNECT4) return launch_t<Connect4>(net, states, list, count_dev, max_n, out, stride, logits_out, stream);
  return launch_t<TicTacToe>(net, states, list, count_dev, max_n, out, stride, logits_out, stream);
}

}  // namespace umma_v1
}  // namespace spb
