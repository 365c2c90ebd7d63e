// engine.cu — tree/game kernels, host engine and the C ABI (include/selfplay_b200.h).
//
// Replaces Mcts::search (ref: src/mcts.rs:196-332), Tree::use_subtree (:161-192) and the self-play
// loop that consumes them (ref: src/learner_concurrent.rs:169-242).  One warp owns one tree; all
// trees of an engine advance in lock-step.  Two search pipelines:
//   * fused  (DetEval / uniform evaluators): ONE kernel runs all `num_searches` simulations of every
//     tree; the path of a simulation lives in registers.
//   * split  (network evaluator): per simulation step  [evaluator kernel] -> [tree_step kernel], where
//     tree_step = expand+backup of the evaluated leaf followed by the select of the next simulation.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "async.cuh"
#include "evaluator.cuh"
#include "tree.cuh"

namespace spb {

#define SPB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return SPB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

constexpr int WARPS_PER_BLOCK = 4;
constexpr int CTL_WORDS = 9 * 32;   // control words of the asynchronous pipeline, one 128-byte line each
constexpr int THREADS = WARPS_PER_BLOCK * 32;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

__global__ void k_reset(Trees T, uint32_t* hist_len, uint8_t* parked, const uint32_t* slots, const PState* roots, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t g = slots ? slots[i] : i;
  PState root = roots ? roots[i] : ps_make(0, 0, 0, 0, SPB_STATUS_ONGOING);
  T.root_state[g] = root;
  T.buf[g] = 0;
  T.live[g] = 1;
  T.n_nodes[g] = 1;
  parked[g] = 0;
  hist_len[g] = 0;   // a restarted slot starts a new trajectory (Tree::with_root_state has empty histories, mcts.rs:86-89)
  NodeRec r;
  r.N = 0; r.W = 0.0f; r.P = 0.0f;
  r.info = make_info(0, 0, ps_status(root));
  T.rec[0][(size_t)g * T.cap] = r;
  T.par[0][(size_t)g * T.cap] = PAR_NONE | (0xFFu << 24);
  for (uint32_t k = 0; k < T.K; ++k) T.leaf_info[g * T.K + k] = 0;
}

// Fused search: all simulations of one tree inside one warp, evaluator in registers.
template <class G, int EVAL>
__global__ void __launch_bounds__(THREADS) k_search_fused(Trees T, uint32_t num_searches) {
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= T.G || !T.live[g]) return;
  const uint32_t b = T.buf[g];
  NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  uint32_t* par = T.par[b] + (size_t)g * T.cap;
  const PState root = T.root_state[g];
  uint32_t n_nodes = T.n_nodes[g];
  unsigned long long ctr[CTR_COUNT] = {0, 0, 0, 0, 0};
  for (uint32_t s = 0; s < num_searches; ++s) {                       // mcts.rs:214
    WarpPath path;
    uint32_t leaf, linfo;
    int depth;
    PState st;
    descend<G>(rec, root, T.c, lane, path, leaf, depth, st, linfo, T.error);
    ctr[CTR_SIMS] += 1;
    ctr[CTR_PATHSUM] += (unsigned)depth;
    const uint32_t status = info_status(linfo);
    float v;
    if (status != SPB_STATUS_ONGOING) {                               // mcts.rs:245-247
      v = terminal_value(status);
      ctr[CTR_TERMINAL] += 1;
    } else {                                                          // mcts.rs:268-284
      float probs[G::A];
      if (EVAL == SPB_EVAL_DET) det_eval<G>(st, probs, &v); else uniform_eval<G>(st, probs, &v);
      ctr[CTR_EVALS] += 1;
      uint32_t before = n_nodes;
      if (!expand<G>(rec, par, T.cap, n_nodes, leaf, st, probs, lane)) {
        if (lane == 0) atomicOr(T.error, ERRBIT_POOL);
        break;
      }
      ctr[CTR_CHILDREN] += n_nodes - before;
    }
    backup_regs(rec, path, depth, v, lane);
    __syncwarp();
  }
  if (lane == 0) T.n_nodes[g] = n_nodes;
  flush_counters(T, ctr, lane);
}

// Split pipeline.  do_finish: expand + backup the leaf whose evaluation is in eval_out.
// do_select: run the select of the next simulation; terminal leaves are backed up at once,
// the others are appended to the evaluator's work list.
#ifdef SPB_TRACE
__device__ unsigned long long g_pdl_trace[64][8];   // trace build: globaltimer stamps of consecutive kernels, [i][0..2] tree step entry / after wait / exit
__device__ unsigned int g_pdl_idx = 0;
__device__ unsigned long long g_warp_trace[8192][6];   // per tree of the latest tree step: entry, after wait, after finish, after descend, after append, exit
extern "C" int spb_debug_warp_trace(unsigned long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, g_warp_trace, sizeof(unsigned long long) * 6 * (size_t)n); }
#define WTRACE(i) do { if (lane == 0 && g < 8192 && do_select && do_finish) g_warp_trace[g][i] = gtimer(); } while (0)
extern "C" int spb_debug_pdl_trace(unsigned long long* out) {
  unsigned int z = 0;
  int rc = (int)cudaMemcpyFromSymbol(out, g_pdl_trace, sizeof(unsigned long long) * 64 * 8);
  rc |= (int)cudaMemcpyToSymbol(g_pdl_idx, &z, sizeof z);
  return rc;
}
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define WTRACE(i) ((void)0)
#endif

template <class G>
__global__ void __launch_bounds__(THREADS) k_tree_step(Trees T, int do_finish, int do_select, uint32_t parity) {
  // Programmatic dependent launch (no-ops for a plain launch): the evaluator that follows may start its set-up while
  // this grid runs, and this grid may have been started before the evaluator in front of it finished.
#ifdef SPB_TRACE
  const unsigned long long tr0 = gtimer();
#endif
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SPB_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned int i = g_pdl_idx++ & 63u;
    g_pdl_trace[i][0] = tr0;
    g_pdl_trace[i][1] = gtimer();
  }
#endif
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) T.eval_count[(parity + 1) & 1] = 0;   // for the NEXT step's select
  if (g >= T.G) return;
  // Everything that is addressed by the tree index alone is requested in ONE round trip (the step is bound by the
  // latency of dependent global loads, not by bandwidth): liveness, live arena, the pending leaf, its state, the
  // evaluator's answer, the arena length and the stored path (lane d holds path node d and d + 32).
  const uint32_t slot = g;   // K == 1
  uint32_t* pathm = T.path + (size_t)slot * G::MAX_DEPTH;
  const uint8_t live = T.live[g];
  const uint32_t b = T.buf[g];
  const uint32_t li = T.leaf_info[slot];
  const PState st = T.leaf_state[slot];
  const float* eo = T.eval_out + (size_t)slot * G::EVAL_STRIDE;
  float probs[G::A];
#pragma unroll
  for (int a = 0; a < G::A; ++a) probs[a] = eo[a];
  const float v = eo[G::A];
  uint32_t n_nodes = T.n_nodes[g];
  const uint32_t pn0 = (lane < G::MAX_DEPTH) ? pathm[lane] : 0u;   // tic-tac-toe paths are 12 words: lanes beyond must not read the next slot's
  const uint32_t pn1 = (lane + 32 < G::MAX_DEPTH) ? pathm[lane + 32] : 0u;
  if (!live) return;
#ifdef SPB_TRACE
  if (lane == 0 && g < 8192 && do_select && do_finish) g_warp_trace[g][0] = tr0;
#endif
  WTRACE(1);
  NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  uint32_t* par = T.par[b] + (size_t)g * T.cap;
  unsigned long long ctr[CTR_COUNT] = {0, 0, 0, 0, 0};
  bool ok = true;

  if (do_finish && (li & LEAF_PENDING)) {
    const int depth = (int)(li & 0xFFu);
    const uint32_t leaf = __shfl_sync(0xffffffffu, depth < 32 ? pn0 : pn1, depth & 31);
    // second round trip: the path nodes' statistics (backup, mcts.rs:145-159), requested before the expand's stores
    uint2 nw0 = make_uint2(0, 0), nw1 = make_uint2(0, 0);
    if (lane <= depth) nw0 = *reinterpret_cast<const uint2*>(&rec[pn0]);
    if (lane + 32 <= depth) nw1 = *reinterpret_cast<const uint2*>(&rec[pn1]);
    const uint32_t before = n_nodes;
    if (!expand<G>(rec, par, T.cap, n_nodes, leaf, st, probs, lane)) {
      if (lane == 0) atomicOr(T.error, ERRBIT_POOL);
      ok = false;
    } else {
      ctr[CTR_CHILDREN] += n_nodes - before;
      if (lane == 0) T.n_nodes[g] = n_nodes;
      if (lane <= depth) {                                        // same arithmetic as backup_mem: N += 1, W = W + (+-v)
        nw0.x += 1u;
        nw0.y = __float_as_uint(__fadd_rn(__uint_as_float(nw0.y), ((depth - lane) & 1) ? -v : v));
        *reinterpret_cast<uint2*>(&rec[pn0]) = nw0;
      }
      if (lane + 32 <= depth) {
        nw1.x += 1u;
        nw1.y = __float_as_uint(__fadd_rn(__uint_as_float(nw1.y), ((depth - lane - 32) & 1) ? -v : v));
        *reinterpret_cast<uint2*>(&rec[pn1]) = nw1;
      }
    }
    if (lane == 0) T.leaf_info[slot] = 0;
    __syncwarp();
  }

  WTRACE(2);
  if (do_select && ok) {
    WarpPath path;
    uint32_t leaf, linfo;
    int depth;
    PState st;
    descend<G>(rec, T.root_state[g], T.c, lane, path, leaf, depth, st, linfo, T.error);
    WTRACE(3);
    ctr[CTR_SIMS] += 1;
    ctr[CTR_PATHSUM] += (unsigned)depth;
    const uint32_t status = info_status(linfo);
    if (status != SPB_STATUS_ONGOING) {
      ctr[CTR_TERMINAL] += 1;
      backup_regs(rec, path, depth, terminal_value(status), lane);
    } else {
      ctr[CTR_EVALS] += 1;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        int d = lane + 32 * s;
        if (d <= depth) pathm[d] = path.node[s];
      }
      if (lane == 0) {
        T.leaf_state[slot] = st;
        T.leaf_info[slot] = (uint32_t)depth | LEAF_PENDING;
        uint32_t pos = atomicAdd(&T.eval_count[parity & 1], 1u);
        T.eval_list[pos] = slot;
      }
    }
  }
  WTRACE(4);
  flush_counters(T, ctr, lane);
  WTRACE(5);
}

// ---- EXTENSION (not in the reference): K in-flight leaves per tree per step with virtual loss ------------
// BASELINE.json config 4 / SURVEY.md §8(f)-3; the reference runs one leaf per tree per step (mcts.rs:236-252)
// and K = 1 never comes here.  The warp that owns the tree runs its K descents one after the other, so no
// atomics are needed and the result is deterministic.  Definition (identical in oracle/oracle.cc search_vl):
//   select k: PUCT descent on the current statistics.  Terminal leaf: real backup at once.  Otherwise every
//   path node takes a virtual loss (N += 1, W += 1.0) and the leaf is queued — as a duplicate if the same leaf
//   is already queued in this step.  Finish (after the evaluator): entries in selection order; a first
//   occurrence expands; every entry rewrites each path node as W = (W - 1.0) + sign*v.
// leaf_info[slot]: depth[0,8) | LEAF_PENDING | LEAF_DUP | source entry[16,24)
constexpr uint32_t LEAF_DUP = 1u << 9;

__device__ __forceinline__ void backup_virtual(NodeRec* rec, const uint32_t* path, int depth, float v, int lane) {
  for (int d = lane; d <= depth; d += 32) {
    const uint32_t node = path[d];
    const float sv = ((depth - d) & 1) ? -v : v;
    float* w = &rec[node].W;
    *w = __fadd_rn(__fsub_rn(*w, 1.0f), sv);
  }
}

template <class G>
__global__ void __launch_bounds__(THREADS) k_tree_step_multi(Trees T, int do_finish, uint32_t k_select, uint32_t parity) {
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) T.eval_count[(parity + 1) & 1] = 0;
  if (g >= T.G || !T.live[g]) return;
  const uint32_t b = T.buf[g];
  NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  uint32_t* par = T.par[b] + (size_t)g * T.cap;
  const uint32_t K = T.K;
  unsigned long long ctr[CTR_COUNT] = {0, 0, 0, 0, 0};
  bool ok = true;

  if (do_finish) {
    uint32_t n_nodes = T.n_nodes[g];
    for (uint32_t k = 0; k < K && ok; ++k) {
      const uint32_t slot = g * K + k;
      const uint32_t li = T.leaf_info[slot];
      if (!(li & (LEAF_PENDING | LEAF_DUP))) continue;
      const int depth = (int)(li & 0xFFu);
      const uint32_t src = (li & LEAF_DUP) ? g * K + ((li >> 16) & 0xFFu) : slot;
      const uint32_t* pathm = T.path + (size_t)src * G::MAX_DEPTH;
      const float* eo = T.eval_out + (size_t)src * G::EVAL_STRIDE;
      if (li & LEAF_PENDING) {
        float probs[G::A];
#pragma unroll
        for (int a = 0; a < G::A; ++a) probs[a] = eo[a];
        const uint32_t before = n_nodes;
        if (!expand<G>(rec, par, T.cap, n_nodes, pathm[depth], T.leaf_state[slot], probs, lane)) {
          if (lane == 0) atomicOr(T.error, ERRBIT_POOL);
          ok = false;
          break;
        }
        ctr[CTR_CHILDREN] += n_nodes - before;
      }
      __syncwarp();
      backup_virtual(rec, pathm, depth, eo[G::A], lane);
      if (lane == 0) T.leaf_info[slot] = 0;
      __syncwarp();
    }
    if (lane == 0) T.n_nodes[g] = n_nodes;
  }

  if (ok) {
    uint32_t my_leaf = 0xFFFFFFFFu;                    // lane j: leaf queued by entry j of this step (first occurrences only)
    const PState root = T.root_state[g];
    for (uint32_t k = 0; k < k_select; ++k) {
      const uint32_t slot = g * K + k;
      WarpPath path;
      uint32_t leaf, linfo;
      int depth;
      PState st;
      descend<G>(rec, root, T.c, lane, path, leaf, depth, st, linfo, T.error);
      ctr[CTR_SIMS] += 1;
      ctr[CTR_PATHSUM] += (unsigned)depth;
      const uint32_t status = info_status(linfo);
      if (status != SPB_STATUS_ONGOING) {
        ctr[CTR_TERMINAL] += 1;
        backup_regs(rec, path, depth, terminal_value(status), lane);
      } else {
        const unsigned dupmask = __ballot_sync(0xffffffffu, my_leaf == leaf);
        // virtual loss on the whole path (lane d <-> depth d)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int d = lane + 32 * s;
          if (d <= depth) {
            uint2 nw;
            nw.x = path.N[s] + 1u;
            nw.y = __float_as_uint(__fadd_rn(path.W[s], 1.0f));
            *reinterpret_cast<uint2*>(&rec[path.node[s]]) = nw;
          }
        }
        if (dupmask) {
          if (lane == 0) T.leaf_info[slot] = (uint32_t)depth | LEAF_DUP | ((uint32_t)(__ffs((int)dupmask) - 1) << 16);
        } else {
          ctr[CTR_EVALS] += 1;
          uint32_t* pathm = T.path + (size_t)slot * G::MAX_DEPTH;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int d = lane + 32 * s;
            if (d <= depth) pathm[d] = path.node[s];
          }
          if (lane == (int)k) my_leaf = leaf;
          if (lane == 0) {
            T.leaf_state[slot] = st;
            T.leaf_info[slot] = (uint32_t)depth | LEAF_PENDING;
            const uint32_t pos = atomicAdd(&T.eval_count[parity & 1], 1u);
            T.eval_list[pos] = slot;
          }
        }
      }
      __syncwarp();
    }
  }
  flush_counters(T, ctr, lane);
}

// Evaluator stand-ins for the split pipeline (parity harness): DetEval / uniform over the work list.
template <class G, int EVAL>
__global__ void k_eval_builtin(const PState* states, const uint32_t* list, const uint32_t* count, float* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *count) return;
  uint32_t slot = list[i];
  float probs[G::A], v;
  if (EVAL == SPB_EVAL_DET) det_eval<G>(states[slot], probs, &v); else uniform_eval<G>(states[slot], probs, &v);
  float* o = out + (size_t)slot * G::EVAL_STRIDE;
#pragma unroll
  for (int a = 0; a < G::A; ++a) o[a] = probs[a];
  o[G::A] = v;
}

// ---- asynchronous pipeline (async.cuh) --------------------------------------------------------------------
// Start of a search: every live tree gets its simulation budget and a ticket of the ready ring.
__global__ void k_async_init(Trees T, AsyncCtl C, uint32_t num_searches) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= T.G || !T.live[g]) return;
  C.sims_left[g] = num_searches;
  atomicAdd(C.n_active, 1u);
  const uint32_t t = atomicAdd(C.ready.tail, 1u);
  C.ready.slots[t & C.ready.mask] = ring_entry(t, g);
}

// The pipeline with the built-in evaluators (parity harness: SPB_FLAG_FORCE_SPLIT): the same rings and tree warps as the
// network pipeline, evaluator CTAs replaced by evaluator warps.  Warps 0,1 evaluate, warps 2,3 own trees.
template <class G, int EVAL>
__global__ void __launch_bounds__(THREADS) k_async_builtin(Trees T, AsyncCtl C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) builtin_eval_worker<G, EVAL>(T, C, gridDim.x * 2u, lane);
  else tree_worker<G>(T, C, lane);
}

// Replays the moves from the root to `node` (walks the parent links up, then down again).
template <class G>
__device__ PState node_state(const Trees& T, uint32_t g, uint32_t node) {
  const uint32_t b = T.buf[g];
  const NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  const uint32_t* par = T.par[b] + (size_t)g * T.cap;
  uint8_t acts[G::MAX_DEPTH];
  int n = 0;
  uint32_t cur = node;
  while (true) {
    uint32_t p = par[cur];
    if ((p & PAR_NONE) == PAR_NONE || n >= G::MAX_DEPTH) break;
    acts[n++] = (uint8_t)(p >> 24);
    cur = p & PAR_NONE;
  }
  PState st = T.root_state[g];
  for (int i = n - 1; i >= 0; --i) st = G::place(st, acts[i], i == 0 ? info_status(rec[node].info) : (uint32_t)SPB_STATUS_ONGOING);
  return st;
}

// use_subtree, mcts.rs:161-192: breadth-first copy of the subtree under `new_root` into the other
// arena.  New ids are BFS order, children stay contiguous and in action order, statistics are kept.
// A window of 32 already-copied nodes is processed per iteration; a warp scan of the child counts
// assigns the children's new ids exactly as the sequential queue would.
template <class G>
__device__ void reroot(const Trees& T, uint32_t g, uint32_t new_root, int lane) {
  const uint32_t b = T.buf[g];
  const NodeRec* orec = T.rec[b] + (size_t)g * T.cap;
  const uint32_t* opar = T.par[b] + (size_t)g * T.cap;
  NodeRec* nrec = T.rec[b ^ 1] + (size_t)g * T.cap;
  uint32_t* npar = T.par[b ^ 1] + (size_t)g * T.cap;
  PState new_state = node_state<G>(T, g, new_root);
  if (lane == 0) {
    nrec[0] = orec[new_root];                                  // info still holds the OLD first_child
    npar[0] = PAR_NONE | (opar[new_root] & 0xFF000000u);       // parent_id = None, action_taken kept (:166-167)
  }
  __syncwarp();
  uint32_t next = 1;
  for (uint32_t lo = 0; lo < next;) {
    const uint32_t hi = min(next, lo + 32u);                   // nodes [lo, hi) are already in the new arena
    const uint32_t i = lo + lane;
    const bool active = i < hi;
    uint32_t info = active ? nrec[i].info : 0u;
    const uint32_t nc = info_nc(info), ofc = info_fc(info);
    uint32_t incl = nc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t nfc = next + incl - nc;
    if (active) {
      for (uint32_t j = 0; j < nc; ++j) {
        nrec[nfc + j] = orec[ofc + j];
        npar[nfc + j] = i | (opar[ofc + j] & 0xFF000000u);
      }
      nrec[i].info = make_info(nc ? nfc : 0u, nc, info_status(info));
    }
    next += total;
    lo = hi;
    __syncwarp();
  }
  if (lane == 0) {
    T.n_nodes[g] = next;
    T.buf[g] = (uint8_t)(b ^ 1);
    T.root_state[g] = new_state;
    for (uint32_t k = 0; k < T.K; ++k) T.leaf_info[g * T.K + k] = 0;   // node_id_to_expand = None (learner_concurrent.rs:233)
  }
}

template <class G>
__global__ void __launch_bounds__(THREADS) k_advance(Trees T, const uint32_t* slots, const uint32_t* node_ids, uint32_t n,
                                                     PState* out_states) {
  const int lane = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint32_t g = slots ? slots[i] : i;
  reroot<G>(T, g, node_ids[i], lane);
  __syncwarp();
  if (lane == 0 && out_states) out_states[i] = T.root_state[g];
}

template <class G>
__global__ void k_get_state(Trees T, uint32_t g, uint32_t node, PState* out) { *out = node_state<G>(T, g, node); }

// Root children of every slot, in child order (mcts.rs:310-331).
__global__ void k_root_children(Trees T, uint8_t* actions, uint32_t* counts, uint32_t* ids, uint32_t* ncs) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= T.G) return;
  const uint32_t b = T.buf[g];
  const NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  const uint32_t* par = T.par[b] + (size_t)g * T.cap;
  uint32_t info = rec[0].info;
  uint32_t nc = T.live[g] ? info_nc(info) : 0u, fc = info_fc(info);
  ncs[g] = nc;
  for (uint32_t j = 0; j < SPB_MAX_ACTIONS; ++j) {
    bool v = j < nc;
    actions[g * SPB_MAX_ACTIONS + j] = v ? (uint8_t)(par[fc + j] >> 24) : (uint8_t)0xFF;
    counts[g * SPB_MAX_ACTIONS + j] = v ? rec[fc + j].N : 0u;
    ids[g * SPB_MAX_ACTIONS + j] = v ? fc + j : 0u;
  }
}

__global__ void k_node_stats(Trees T, uint32_t g, uint32_t node, NodeRec* out) {
  *out = T.rec[T.buf[g]][(size_t)g * T.cap + node];
}

__global__ void k_nodes_live(Trees T, unsigned long long* out) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < T.G && T.live[g]) atomicAdd(out, (unsigned long long)T.n_nodes[g]);
}

__global__ void k_max_nodes(Trees T, unsigned long long* out) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < T.G && T.live[g]) atomicMax(out, (unsigned long long)T.n_nodes[g]);
}

// ---- State trait, batched (ref: game/mod.rs:21-33) ------------------------------------------------
template <class G>
__global__ void k_game_next(const PState* in, const uint8_t* actions, uint32_t n, PState* out, int32_t* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PState o = in[i];
  bool ok = G::next_state(in[i], actions[i], &o);
  out[i] = o;
  err[i] = ok ? SPB_OK : SPB_ERR_ILLEGAL;
}
template <class G>
__global__ void k_game_valid(const PState* in, uint32_t n, uint32_t* masks) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) masks[i] = G::valid_mask(in[i]);
}
template <class G>
__global__ void k_game_encode(const PState* in, uint32_t n, float* out) {
  constexpr int E = 3 * G::ROWS * G::COLS;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * E) return;
  uint32_t i = (uint32_t)(t / E);
  int r = (int)(t % E);
  int plane = r / (G::ROWS * G::COLS), row = (r / G::COLS) % G::ROWS, col = r % G::COLS;
  out[t] = G::encode_cell(in[i], plane, row, col);
}

// Masks + renormalises evaluator output for spb_predict (model/mod.rs:86-93).
template <class G>
__global__ void k_mask_policies(const PState* states, uint32_t n, const float* eval_out, float* policies, float* values) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float probs[G::A], pri[G::A];
#pragma unroll
  for (int a = 0; a < G::A; ++a) probs[a] = eval_out[(size_t)i * G::EVAL_STRIDE + a];
  mask_renorm<G>(G::valid_mask(states[i]), probs, pri);
#pragma unroll
  for (int a = 0; a < G::A; ++a) policies[(size_t)i * G::A + a] = pri[a];
  values[i] = eval_out[(size_t)i * G::EVAL_STRIDE + G::A];
}

// ---- self-play ply (ref: learner_concurrent.rs:179-238) -------------------------------------------
struct SelfPlay {
  spb_position* hist;        // [G][MAX_PLY]
  uint32_t* hist_len;        // [G]
  uint8_t* parked;           // [G]  0, or 1 | terminal status << 1 | terminal side to move << 3: the game has ended but its trajectory
                             //      did not fit the output buffer; the slot is idle until the next spb_selfplay_step emits it
  unsigned long long* game_id;     // [G]
  spb_position* out;         // [out_cap]
  unsigned long long* out_game;    // [out_cap]
  unsigned long long* out_cursor;  // [1]
  uint32_t* finished;        // [1]
  uint32_t out_cap;
  uint32_t max_ply;
  unsigned long long id_stride;
};
enum : uint32_t { ERRBIT_TRAJ_FULL = 4u };

// Reserves `plies` records of the trajectory output buffer (lane 0; all lanes get the answer).  The cursor only moves
// when the whole trajectory fits, so every record below the cursor is fully written: a drain never sees a hole.
__device__ __forceinline__ bool reserve_output(const SelfPlay& P, uint32_t plies, int lane, unsigned long long* base_out) {
  unsigned long long base = ~0ull;
  if (lane == 0) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(P.out_cursor);
    for (;;) {
      if (cur + plies > P.out_cap) { base = ~0ull; break; }
      const unsigned long long seen = atomicCAS(P.out_cursor, cur, cur + plies);
      if (seen == cur) { base = cur; break; }
      cur = seen;
    }
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  *base_out = base;
  return base != ~0ull;
}

// Emits the finished game of slot g (learner_concurrent.rs:200-230) and restarts or retires the slot.
__device__ __forceinline__ void emit_and_finish(const Trees& T, const SelfPlay& P, uint32_t g, uint32_t plies, uint32_t cstatus,
                                                uint32_t term_player, unsigned long long base, const PState* restart_roots, int lane) {
  const float value = terminal_value(cstatus);                     // from the terminal state's side to move
  for (uint32_t i = lane; i < plies; i += 32) {
    spb_position p = P.hist[(size_t)g * P.max_ply + i];
    float v = (p.current_player == term_player) ? value : -value;   // :214-226
    p.outcome = (int8_t)v;
    P.out[base + i] = p;
    P.out_game[base + i] = P.game_id[g];
  }
  __syncwarp();
  if (lane == 0) {
    atomicAdd(P.finished, 1u);
    P.hist_len[g] = 0;
    P.parked[g] = 0;
    P.game_id[g] += P.id_stride;
    if (restart_roots) {                                           // Tree::with_root_state for the next game
      PState nr = restart_roots[g];
      T.root_state[g] = nr;
      T.buf[g] = 0;
      T.n_nodes[g] = 1;
      T.live[g] = 1;
      NodeRec r; r.N = 0; r.W = 0.0f; r.P = 0.0f; r.info = make_info(0, 0, ps_status(nr));
      T.rec[0][(size_t)g * T.cap] = r;
      T.par[0][(size_t)g * T.cap] = PAR_NONE | (0xFFu << 24);
      for (uint32_t k = 0; k < T.K; ++k) T.leaf_info[g * T.K + k] = 0;
    } else {
      T.live[g] = 0;                                               // trees_vec.remove(i), :230
    }
  }
}

template <class G>
__global__ void __launch_bounds__(THREADS) k_selfplay_step(Trees T, SelfPlay P, int rule, float temperature,
                                                           unsigned long long seed, const PState* restart_roots) {
  const int lane = threadIdx.x & 31;
  const uint32_t g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= T.G) return;
  const uint32_t pk = P.parked[g];
  if (pk) {
    // A game that ended in an earlier call while the output buffer was full: the slot has been idle since; emit now.
    const uint32_t plies = P.hist_len[g];
    unsigned long long base;
    if (reserve_output(P, plies, lane, &base)) emit_and_finish(T, P, g, plies, (pk >> 1) & 3u, (pk >> 3) & 1u, base, restart_roots, lane);
    else if (lane == 0) atomicOr(T.error, ERRBIT_TRAJ_FULL);
    return;
  }
  if (!T.live[g]) return;
  const uint32_t b = T.buf[g];
  NodeRec* rec = T.rec[b] + (size_t)g * T.cap;
  const uint32_t* par = T.par[b] + (size_t)g * T.cap;
  const uint32_t info = rec[0].info;
  const uint32_t nc = info_nc(info), fc = info_fc(info);
  if (nc == 0) return;                                           // nothing searched / terminal root
  const PState root = T.root_state[g];
  uint32_t cnt = 0, act = 0, cinfo = 0;
  if (lane < (int)nc) { cnt = rec[fc + lane].N; cinfo = rec[fc + lane].info; act = par[fc + lane] >> 24; }
  int chosen;
  if (rule == SPB_MOVE_TEMPERATURE) {
    // learner_concurrent.rs:189-194: WeightedIndex over count^temperature.  Counter-based RNG (the
    // reference uses the unseedable thread_rng, so there is no stream to match).
    float w = lane < (int)nc ? powf((float)cnt, temperature) : 0.0f;
    float incl = w;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    float total = __shfl_sync(0xffffffffu, incl, 15);
    unsigned long long h = splitmix64(seed ^ splitmix64(P.game_id[g] * 64ull + P.hist_len[g]));
    float u = (float)(h >> 40) * (1.0f / 16777216.0f) * total;
    unsigned ball = __ballot_sync(0xffffffffu, lane < (int)nc && u < incl && w > 0.0f);
    chosen = ball ? __ffs((int)ball) - 1 : (int)nc - 1;
  } else {
    // main.rs:108-112: max_by(total_cmp) over visit counts -> the LAST maximal child.
    chosen = warp_argmax_last<G::A>(lane < (int)nc ? (float)cnt : -INFINITY, lane < (int)nc ? lane : -1);
    chosen = __shfl_sync(0xffffffffu, chosen, 0);
  }
  // learner_concurrent.rs:197-198: push root state + visit-count policy.
  const uint32_t ply = P.hist_len[g];
  if (ply < P.max_ply) {
    spb_position* h = &P.hist[(size_t)g * P.max_ply + ply];
    if (lane == 0) {
      h->stones[0] = ps_x(root); h->stones[1] = ps_o(root);
      h->current_player = (uint8_t)ps_player(root);
      h->ply = (uint8_t)ply; h->outcome = 0; h->reserved = 0;
      for (int a = 0; a < SPB_MAX_ACTIONS; ++a) h->visit_counts[a] = 0;
    }
    __syncwarp();
    if (lane < (int)nc && act < SPB_MAX_ACTIONS) h->visit_counts[act] = cnt;
    __syncwarp();
  }
  const uint32_t cstatus = info_status(__shfl_sync(0xffffffffu, cinfo, chosen));
  const uint32_t plies = min(ply + 1, P.max_ply);
  if (cstatus != SPB_STATUS_ONGOING) {
    // learner_concurrent.rs:200-230: the game is over.  The trajectory is emitted only when all of it fits the output
    // buffer; otherwise the slot is parked (idle, history kept) and the call reports SPB_ERR_STATE: the caller drains
    // and the next spb_selfplay_step emits the parked games, so no game and no record is ever lost or half-written.
    const uint32_t term_player = ps_player(root) ^ 1u;
    unsigned long long base;
    if (reserve_output(P, plies, lane, &base)) {
      emit_and_finish(T, P, g, plies, cstatus, term_player, base, restart_roots, lane);
    } else if (lane == 0) {
      P.hist_len[g] = plies;
      P.parked[g] = (uint8_t)(1u | (cstatus << 1) | (term_player << 3));
      T.live[g] = 0;
      atomicOr(T.error, ERRBIT_TRAJ_FULL);
    }
  } else {
    if (lane == 0) P.hist_len[g] = plies;
    reroot<G>(T, g, fc + (uint32_t)chosen, lane);                // :233-234
  }
}

}  // namespace spb

// ------------------------------------------------------------------------------------------------
// host engine
// ------------------------------------------------------------------------------------------------
using namespace spb;

static thread_local std::string g_create_error;

struct spb_engine {
  spb_config cfg{};
  std::string err;
  Trees T{};
  SelfPlay P{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<void*> allocs;
  Evaluator evaluator;
  // staging
  void* h_stage = nullptr; size_t h_stage_bytes = 0;   // pinned
  void* d_stage = nullptr; size_t d_stage_bytes = 0;
  uint8_t* d_rc_actions = nullptr; uint32_t* d_rc_counts = nullptr; uint32_t* d_rc_ids = nullptr; uint32_t* d_rc_n = nullptr;
  unsigned long long* d_misc = nullptr;   // [4] scratch u64
  uint64_t launches = 0;
  float last_search_ms = 0.f, last_eval_ms = 0.f;
  uint32_t last_eval_launches = 0;
  int last_eval_parity = -1;   // parity of the work list the last evaluator launch of spb_search consumed
  // CUDA graph of one split-pipeline step pair (parity 0 and 1)
  cudaGraphExec_t step_graph = nullptr;
  // asynchronous pipeline: rings + control words (async.cuh)
  AsyncCtl ctl{};
  uint32_t* d_ctl_words = nullptr;   // [CTL_WORDS] one 128-byte line per counter
  unsigned long long* d_ring_slots = nullptr;   // [2][ring_size]
  uint32_t ring_size = 0;
  uint64_t last_async_stats[16] = {};
  int A = 0, max_depth = 0, eval_stride = 0, max_ply = 0;

  void set_error(const std::string& s) { err = s; }

  template <class T_> int32_t dalloc(T_** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T_));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return SPB_ERR_NOMEM; }
    allocs.push_back(q);
    *p = static_cast<T_*>(q);
    return SPB_OK;
  }
  int32_t ensure_capacity(uint32_t num_searches);
  int32_t ensure_stage(size_t bytes) {
    if (bytes > h_stage_bytes) {
      if (h_stage) cudaFreeHost(h_stage);
      h_stage = nullptr; h_stage_bytes = 0;
      size_t nb = std::max(bytes, (size_t)1 << 20);
      SPB_CUDA(cudaMallocHost(&h_stage, nb));
      h_stage_bytes = nb;
    }
    if (bytes > d_stage_bytes) {
      if (d_stage) cudaFree(d_stage);
      d_stage = nullptr; d_stage_bytes = 0;
      size_t nb = std::max(bytes, (size_t)1 << 20);
      SPB_CUDA(cudaMalloc(&d_stage, nb));
      d_stage_bytes = nb;
    }
    return SPB_OK;
  }
  int32_t check_device_errors();
  int32_t init();
  void destroy();
  template <class G> int32_t search_t(uint32_t num_searches);
  template <class G> int32_t launch_eval_step(uint32_t parity);
  template <class G> int32_t search_async(uint32_t num_searches);
};

#define SPB_CHECK_LAUNCH() SPB_CUDA(cudaGetLastError())

int32_t spb_engine::init() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (this library has no CPU fallback)"); return SPB_ERR_CUDA; }
  if (cfg.device < 0 || cfg.device >= ndev) { set_error("device ordinal out of range"); return SPB_ERR_ARG; }
  SPB_CUDA(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  SPB_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
  if (prop.major != 10) { set_error("device is not sm_100 (B200); this library is built for sm_100a only"); return SPB_ERR_CUDA; }
  SPB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  SPB_CUDA(cudaEventCreate(&ev0));
  SPB_CUDA(cudaEventCreate(&ev1));
  const bool c4 = cfg.game == SPB_GAME_CONNECT4;
  A = c4 ? Connect4::A : TicTacToe::A;
  max_depth = c4 ? Connect4::MAX_DEPTH : TicTacToe::MAX_DEPTH;
  eval_stride = c4 ? Connect4::EVAL_STRIDE : TicTacToe::EVAL_STRIDE;
  max_ply = c4 ? 42 : 9;
  T.G = cfg.num_games; T.K = cfg.leaves_per_tree; T.cap = cfg.max_nodes_per_tree; T.c = cfg.c;
  const size_t G = T.G, slots = G * T.K, pool = G * (size_t)T.cap;
  int32_t rc;
  for (int b = 0; b < 2; ++b) {
    if ((rc = dalloc(&T.rec[b], pool))) return rc;
    if ((rc = dalloc(&T.par[b], pool))) return rc;
  }
  if ((rc = dalloc(&T.root_state, G))) return rc;
  if ((rc = dalloc(&T.n_nodes, G))) return rc;
  if ((rc = dalloc(&T.buf, G))) return rc;
  if ((rc = dalloc(&T.live, G))) return rc;
  if ((rc = dalloc(&T.path, slots * max_depth))) return rc;
  if ((rc = dalloc(&T.leaf_info, slots))) return rc;
  if ((rc = dalloc(&T.leaf_state, slots))) return rc;
  if ((rc = dalloc(&T.eval_list, slots))) return rc;
  if ((rc = dalloc(&T.eval_count, 4))) return rc;   // [0],[1] ping-pong, [2] snapshot for spb_time_evaluator
  if ((rc = dalloc(&T.eval_out, slots * eval_stride))) return rc;
  if ((rc = dalloc(&T.counters, (size_t)CTR_COUNT))) return rc;
  if ((rc = dalloc(&T.error, 1))) return rc;
  if ((rc = dalloc(&d_rc_actions, G * SPB_MAX_ACTIONS))) return rc;
  if ((rc = dalloc(&d_rc_counts, G * SPB_MAX_ACTIONS))) return rc;
  if ((rc = dalloc(&d_rc_ids, G * SPB_MAX_ACTIONS))) return rc;
  if ((rc = dalloc(&d_rc_n, G))) return rc;
  if ((rc = dalloc(&d_misc, 4))) return rc;
  // asynchronous pipeline: two rings of >= 2G entries (a ring never holds more than G) and the control words
  {
    ring_size = 1024;
    while (ring_size < 2 * (size_t)G) ring_size <<= 1;
    if ((rc = dalloc(&d_ring_slots, 2 * (size_t)ring_size))) return rc;
    if ((rc = dalloc(&d_ctl_words, (size_t)CTL_WORDS))) return rc;
    if ((rc = dalloc(&ctl.sims_left, G))) return rc;
    ctl.leaf.slots = d_ring_slots;            ctl.leaf.mask = ring_size - 1;
    ctl.ready.slots = d_ring_slots + ring_size; ctl.ready.mask = ring_size - 1;
    ctl.leaf.head = d_ctl_words + 0 * 32;  ctl.leaf.tail = d_ctl_words + 1 * 32;
    ctl.ready.head = d_ctl_words + 2 * 32; ctl.ready.tail = d_ctl_words + 3 * 32;
    ctl.n_active = d_ctl_words + 4 * 32;   ctl.done_count = d_ctl_words + 5 * 32;
    ctl.abort = d_ctl_words + 6 * 32;
    ctl.stats = reinterpret_cast<unsigned long long*>(d_ctl_words + 7 * 32);   // 2 lines: ASTAT_COUNT u64
    ctl.stall_ns = 2000000000ull;             // 2 s without a single hand-off anywhere: a bug, reported as SPB_ERR_STATE
    SPB_CUDA(cudaMemsetAsync(ctl.sims_left, 0, G * 4, stream));
  }
  // self-play buffers
  P.max_ply = (uint32_t)max_ply;
  P.out_cap = (uint32_t)std::min<size_t>(G * max_ply * 4, (size_t)1 << 26);
  if (cfg.trajectory_capacity) P.out_cap = std::max<uint32_t>(cfg.trajectory_capacity, (uint32_t)max_ply);   // a whole game always fits
  P.id_stride = cfg.game_id_stride ? cfg.game_id_stride : (unsigned long long)G;
  if ((rc = dalloc(&P.hist, G * max_ply))) return rc;
  if ((rc = dalloc(&P.hist_len, G))) return rc;
  if ((rc = dalloc(&P.parked, G))) return rc;
  if ((rc = dalloc(&P.game_id, G))) return rc;
  if ((rc = dalloc(&P.out, (size_t)P.out_cap))) return rc;
  if ((rc = dalloc(&P.out_game, (size_t)P.out_cap))) return rc;
  if ((rc = dalloc(&P.out_cursor, 1))) return rc;
  if ((rc = dalloc(&P.finished, 1))) return rc;
  SPB_CUDA(cudaMemsetAsync(T.live, 0, G, stream));
  SPB_CUDA(cudaMemsetAsync(T.buf, 0, G, stream));
  SPB_CUDA(cudaMemsetAsync(T.n_nodes, 0, G * 4, stream));
  SPB_CUDA(cudaMemsetAsync(T.leaf_info, 0, slots * 4, stream));
  SPB_CUDA(cudaMemsetAsync(T.leaf_state, 0, slots * sizeof(PState), stream));
  SPB_CUDA(cudaMemsetAsync(T.eval_count, 0, 16, stream));
  SPB_CUDA(cudaMemsetAsync(T.eval_out, 0, slots * eval_stride * 4, stream));
  SPB_CUDA(cudaMemsetAsync(T.counters, 0, CTR_COUNT * 8, stream));
  SPB_CUDA(cudaMemsetAsync(T.error, 0, 4, stream));
  SPB_CUDA(cudaMemsetAsync(P.hist_len, 0, G * 4, stream));
  SPB_CUDA(cudaMemsetAsync(P.parked, 0, G, stream));
  SPB_CUDA(cudaMemsetAsync(P.out_cursor, 0, 8, stream));
  SPB_CUDA(cudaMemsetAsync(P.finished, 0, 4, stream));
  {
    std::vector<unsigned long long> ids(G);
    unsigned long long base = cfg.game_id_base;
    for (size_t i = 0; i < G; ++i) ids[i] = base + i;
    SPB_CUDA(cudaMemcpyAsync(P.game_id, ids.data(), G * 8, cudaMemcpyHostToDevice, stream));
    SPB_CUDA(cudaStreamSynchronize(stream));
  }
  return SPB_OK;
}

void spb_engine::destroy() {
  if (step_graph) cudaGraphExecDestroy(step_graph);
  for (void* p : allocs) cudaFree(p);
  allocs.clear();
  if (h_stage) cudaFreeHost(h_stage);
  if (d_stage) cudaFree(d_stage);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (stream) cudaStreamDestroy(stream);
}

// The reference's arena is a Vec that grows without bound (mcts.rs:19,151-157).  Here every tree owns a fixed
// slice of the pools, so before a search the slices are widened if the largest live tree could outgrow them:
// one simulation appends at most A children (mcts.rs:131-158).  Growth is a strided device copy; it happens
// before any tree is touched, so a failure leaves the engine as it was.
int32_t spb_engine::ensure_capacity(uint32_t num_searches) {
  unsigned long long mx = 0;
  SPB_CUDA(cudaMemsetAsync(d_misc, 0, 8, stream));
  k_max_nodes<<<(T.G + 127) / 128, 128, 0, stream>>>(T, d_misc);
  SPB_CHECK_LAUNCH();
  ++launches;
  SPB_CUDA(cudaMemcpyAsync(&mx, d_misc, 8, cudaMemcpyDeviceToHost, stream));
  SPB_CUDA(cudaStreamSynchronize(stream));
  const unsigned long long need = mx + (unsigned long long)num_searches * (unsigned long long)A;
  if (need <= T.cap || (cfg.flags & SPB_FLAG_FIXED_POOL)) return SPB_OK;   // fixed pool: the device flag reports exhaustion
  unsigned long long ncap = std::max<unsigned long long>(need, 2ull * T.cap);
  ncap = (ncap + 1023ull) & ~1023ull;
  if (ncap > MAX_CAP) ncap = MAX_CAP;
  if (need > ncap) { set_error("node pool exhausted: a tree would exceed 2^24 nodes"); return SPB_ERR_POOL; }
  const size_t pool = (size_t)T.G * (size_t)ncap;
  NodeRec* nrec[2] = {nullptr, nullptr};
  uint32_t* npar[2] = {nullptr, nullptr};
  cudaError_t ce = cudaSuccess;
  for (int b = 0; b < 2 && ce == cudaSuccess; ++b) {
    ce = cudaMalloc((void**)&nrec[b], pool * sizeof(NodeRec));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&npar[b], pool * sizeof(uint32_t));
  }
  for (int b = 0; b < 2 && ce == cudaSuccess; ++b) {
    ce = cudaMemcpy2DAsync(nrec[b], ncap * sizeof(NodeRec), T.rec[b], (size_t)T.cap * sizeof(NodeRec), (size_t)T.cap * sizeof(NodeRec), T.G,
                           cudaMemcpyDeviceToDevice, stream);
    if (ce == cudaSuccess)
      ce = cudaMemcpy2DAsync(npar[b], ncap * sizeof(uint32_t), T.par[b], (size_t)T.cap * sizeof(uint32_t), (size_t)T.cap * sizeof(uint32_t), T.G,
                             cudaMemcpyDeviceToDevice, stream);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
  if (ce != cudaSuccess) {
    for (int b = 0; b < 2; ++b) { if (nrec[b]) cudaFree(nrec[b]); if (npar[b]) cudaFree(npar[b]); }
    cudaGetLastError();
    set_error(std::string("node pool exhausted: cannot grow the pools to ") + std::to_string(ncap) + " nodes per tree (" + cudaGetErrorString(ce) + ")");
    return SPB_ERR_POOL;
  }
  for (int b = 0; b < 2; ++b) {
    for (void* old : {(void*)T.rec[b], (void*)T.par[b]}) {
      allocs.erase(std::find(allocs.begin(), allocs.end(), old));
      cudaFree(old);
    }
    T.rec[b] = nrec[b]; T.par[b] = npar[b];
    allocs.push_back(nrec[b]); allocs.push_back(npar[b]);
  }
  T.cap = (uint32_t)ncap;
  cfg.max_nodes_per_tree = T.cap;
  if (step_graph) { cudaGraphExecDestroy(step_graph); step_graph = nullptr; }   // the graph captured the old pointers
  return SPB_OK;
}

int32_t spb_engine::check_device_errors() {
  uint32_t e = 0;
  SPB_CUDA(cudaMemcpyAsync(&e, T.error, 4, cudaMemcpyDeviceToHost, stream));
  SPB_CUDA(cudaStreamSynchronize(stream));
  if (e) {
    SPB_CUDA(cudaMemsetAsync(T.error, 0, 4, stream));
    if (e & ERRBIT_POOL) { set_error("node pool exhausted: raise spb_config.max_nodes_per_tree"); return SPB_ERR_POOL; }
    if (e & ERRBIT_TRAJ_FULL) { set_error("trajectory buffer full: call spb_drain_trajectories"); return SPB_ERR_STATE; }
    if (e & ERRBIT_NAN) { set_error("NaN PUCT score (the reference panics here, mcts.rs:109)"); return SPB_ERR_STATE; }
  }
  return SPB_OK;
}

template <class G>
int32_t spb_engine::launch_eval_step(uint32_t parity) {
  const uint32_t slots = T.G * T.K;
  if (cfg.evaluator == SPB_EVAL_NET) {
    cudaError_t e = evaluator.launch(T.leaf_state, T.eval_list, &T.eval_count[parity & 1], slots, T.eval_out, G::EVAL_STRIDE,
                                     nullptr, (cfg.flags & SPB_FLAG_EVAL_SIMT) != 0, stream);
    if (e != cudaSuccess) { set_error(std::string("evaluator launch: ") + cudaGetErrorString(e)); return SPB_ERR_CUDA; }
  } else if (cfg.evaluator == SPB_EVAL_DET) {
    k_eval_builtin<G, SPB_EVAL_DET><<<(slots + 255) / 256, 256, 0, stream>>>(T.leaf_state, T.eval_list, &T.eval_count[parity & 1], T.eval_out);
  } else {
    k_eval_builtin<G, SPB_EVAL_UNIFORM><<<(slots + 255) / 256, 256, 0, stream>>>(T.leaf_state, T.eval_list, &T.eval_count[parity & 1], T.eval_out);
  }
  ++launches;
  return SPB_OK;
}

// Launch with programmatic stream serialization: the grid may begin before its predecessor in the stream has drained
// (it blocks in griddepcontrol.wait before touching anything the predecessor writes).  Only for kernels that contain
// that wait.
template <class... P, class... A>
static cudaError_t launch_pdl(void (*kernel)(P...), uint32_t blocks, uint32_t threads, cudaStream_t stream, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

template <class G>
int32_t spb_engine::search_t(uint32_t num_searches) {
  if (num_searches == 0) return SPB_OK;
  const uint32_t blocks = (T.G + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  const bool split = cfg.evaluator == SPB_EVAL_NET || (cfg.flags & SPB_FLAG_FORCE_SPLIT) || T.K > 1;
  if (cfg.evaluator == SPB_EVAL_NET && !evaluator.loaded()) { set_error("no weights loaded: call spb_load_weights first"); return SPB_ERR_STATE; }
  { int32_t rc = ensure_capacity(num_searches); if (rc != SPB_OK) return rc; }
  // The network evaluator (and the built-in evaluators under SPB_FLAG_FORCE_SPLIT) run through the asynchronous
  // pipeline unless the lock-step pipeline is asked for; K > 1 leaves per tree is a lock-step extension.
  const bool async = split && T.K == 1 && !(cfg.flags & SPB_FLAG_LOCKSTEP);
  SPB_CUDA(cudaEventRecord(ev0, stream));
  last_eval_launches = 0;
  if (async) {
    int32_t rc = search_async<G>(num_searches);
    if (rc != SPB_OK) return rc;
  } else if (!split) {
    if (cfg.evaluator == SPB_EVAL_DET) k_search_fused<G, SPB_EVAL_DET><<<blocks, THREADS, 0, stream>>>(T, num_searches);
    else k_search_fused<G, SPB_EVAL_UNIFORM><<<blocks, THREADS, 0, stream>>>(T, num_searches);
    SPB_CHECK_LAUNCH();
    ++launches;
  } else if (T.K > 1) {
    // EXTENSION: K leaves per tree per step with virtual loss (see k_tree_step_multi).
    const uint32_t K = T.K, steps = (num_searches + K - 1) / K;
    SPB_CUDA(cudaMemsetAsync(T.eval_count, 0, 8, stream));
    k_tree_step_multi<G><<<blocks, THREADS, 0, stream>>>(T, 0, std::min(K, num_searches), 0u);
    SPB_CHECK_LAUNCH();
    ++launches;
    for (uint32_t j = 0; j < steps; ++j) {
      const bool last = (j == steps - 1);
      if (last) {
        SPB_CUDA(cudaMemcpyAsync(&T.eval_count[2], &T.eval_count[j & 1u], 4, cudaMemcpyDeviceToDevice, stream));
        last_eval_parity = 2;
      }
      int32_t rc = launch_eval_step<G>(j);
      if (rc != SPB_OK) return rc;
      SPB_CHECK_LAUNCH();
      ++last_eval_launches;
      const uint32_t done_after = std::min(num_searches, (j + 1) * K);
      const uint32_t k_next = last ? 0u : std::min(K, num_searches - done_after);
      k_tree_step_multi<G><<<blocks, THREADS, 0, stream>>>(T, 1, k_next, j + 1);
      SPB_CHECK_LAUNCH();
      ++launches;
    }
  } else {
    // step j: select writes eval_count[j&1]; the kernel also zeroes eval_count[(j+1)&1].
    SPB_CUDA(cudaMemsetAsync(T.eval_count, 0, 8, stream));
    SPB_CUDA(launch_pdl(k_tree_step<G>, blocks, THREADS, stream, T, 0, 1, 0u));
    ++launches;
    uint32_t j = 0;
    const bool use_graph = !(cfg.flags & SPB_FLAG_NO_GRAPH) && num_searches > 2;
    if (use_graph) {
      // One graph = two steps (parities 0,1): eval(0) step(1) eval(1) step(0).
      if (!step_graph) {
        cudaGraph_t graph;
        SPB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int32_t rc = launch_eval_step<G>(0u);
        cudaError_t pe = launch_pdl(k_tree_step<G>, blocks, THREADS, stream, T, 1, 1, 1u);
        if (rc == SPB_OK) rc = launch_eval_step<G>(1u);
        if (pe == cudaSuccess) pe = launch_pdl(k_tree_step<G>, blocks, THREADS, stream, T, 1, 1, 0u);
        cudaError_t ce = cudaStreamEndCapture(stream, &graph);
        launches -= 2;
        if (rc != SPB_OK) return rc;
        if (ce == cudaSuccess) ce = pe;
        if (ce != cudaSuccess) { set_error(std::string("graph capture: ") + cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
        SPB_CUDA(cudaGraphInstantiate(&step_graph, graph, 0));
        cudaGraphDestroy(graph);
      }
      while (j + 2 <= num_searches - 1) {
        SPB_CUDA(cudaGraphLaunch(step_graph, stream));
        launches += 4;
        last_eval_launches += 2;
        j += 2;
      }
    }
    for (; j < num_searches; ++j) {
      int32_t rc = launch_eval_step<G>(j);
      if (rc != SPB_OK) return rc;
      SPB_CHECK_LAUNCH();
      ++last_eval_launches;
      const int last = (j == num_searches - 1);
      if (last) {   // keep the size of the last work list: the final tree step clears the ping-pong counter
        SPB_CUDA(cudaMemcpyAsync(&T.eval_count[2], &T.eval_count[j & 1u], 4, cudaMemcpyDeviceToDevice, stream));
        last_eval_parity = 2;
      }
      SPB_CUDA(launch_pdl(k_tree_step<G>, blocks, THREADS, stream, T, 1, last ? 0 : 1, j + 1));
      ++launches;
    }
  }
  SPB_CUDA(cudaEventRecord(ev1, stream));
  uint32_t ctl_host[CTL_WORDS];
  if (async) SPB_CUDA(cudaMemcpyAsync(ctl_host, d_ctl_words, sizeof ctl_host, cudaMemcpyDeviceToHost, stream));
  int32_t rc = check_device_errors();   // synchronises the stream
  SPB_CUDA(cudaEventElapsedTime(&last_search_ms, ev0, ev1));
  if (async) std::memcpy(last_async_stats, &ctl_host[7 * 32], sizeof(uint64_t) * ASTAT_COUNT);
  if (rc == SPB_OK && async && (ctl_host[6 * 32] != 0u || ctl_host[5 * 32] != ctl_host[4 * 32])) {
    char msg[256];
    std::snprintf(msg, sizeof msg, "asynchronous search pipeline stalled (abort=%u, trees done %u of %u, leaf ring %u/%u, ready ring %u/%u)",
                  ctl_host[6 * 32], ctl_host[5 * 32], ctl_host[4 * 32], ctl_host[0], ctl_host[32], ctl_host[64], ctl_host[96]);
    set_error(msg);
    return SPB_ERR_STATE;
  }
  return rc;
}

// One search through the asynchronous pipeline: rings reset, every live tree queued, ONE resident kernel until all trees
// have completed their simulations (network: evaluator CTAs + tree warps; built-in evaluators: evaluator warps + tree warps).
template <class G>
int32_t spb_engine::search_async(uint32_t num_searches) {
  SPB_CUDA(cudaMemsetAsync(d_ctl_words, 0, (size_t)CTL_WORDS * 4, stream));
  SPB_CUDA(cudaMemsetAsync(d_ring_slots, 0, 2 * (size_t)ring_size * 8, stream));
  k_async_init<<<(T.G + 127) / 128, 128, 0, stream>>>(T, ctl, num_searches);
  SPB_CHECK_LAUNCH();
  ++launches;
  if (cfg.evaluator == SPB_EVAL_NET) {
    cudaError_t e = evaluator.launch_ring(T, ctl, stream);
    if (e != cudaSuccess) { set_error(std::string("evaluator launch: ") + cudaGetErrorString(e)); return SPB_ERR_CUDA; }
    ++last_eval_launches;
  } else {
    const uint32_t grid = std::max(1u, std::min(592u, (T.G + 3u) / 4u));
    if (cfg.evaluator == SPB_EVAL_DET) k_async_builtin<G, SPB_EVAL_DET><<<grid, THREADS, 0, stream>>>(T, ctl);
    else k_async_builtin<G, SPB_EVAL_UNIFORM><<<grid, THREADS, 0, stream>>>(T, ctl);
    SPB_CHECK_LAUNCH();
  }
  ++launches;
  last_eval_parity = -1;
  return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define ENGINE_GUARD(e)                                   \
  if (!(e)) return SPB_ERR_ARG;                           \
  if (cudaSetDevice((e)->cfg.device) != cudaSuccess) {    \
    (e)->set_error("cudaSetDevice failed");               \
    return SPB_ERR_CUDA;                                  \
  }
#define ARG_CHECK(e, cond, msg)        \
  if (!(cond)) {                       \
    (e)->set_error(msg);               \
    return SPB_ERR_ARG;                \
  }
#define DISPATCH_GAME(e, CALL_C4, CALL_TTT) ((e)->cfg.game == SPB_GAME_CONNECT4 ? (CALL_C4) : (CALL_TTT))

extern "C" {

int32_t spb_abi_version(void) { return SPB_ABI_VERSION; }

int32_t spb_default_config(spb_config* cfg) {
  if (!cfg) return SPB_ERR_ARG;
  std::memset(cfg, 0, sizeof *cfg);
  cfg->abi_version = SPB_ABI_VERSION;
  cfg->game = SPB_GAME_CONNECT4;
  cfg->device = 0;
  cfg->num_games = 100;            // mcts.rs:54
  cfg->max_nodes_per_tree = 0;
  cfg->leaves_per_tree = 1;
  cfg->c = 2.0f;                   // mcts.rs:49
  cfg->evaluator = SPB_EVAL_NET;
  return SPB_OK;
}

int32_t spb_create(const spb_config* cfg, spb_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return SPB_ERR_ARG; }
  *out = nullptr;
  if (cfg->abi_version != SPB_ABI_VERSION) { g_create_error = "abi_version mismatch"; return SPB_ERR_ARG; }
  if (cfg->game != SPB_GAME_CONNECT4 && cfg->game != SPB_GAME_TICTACTOE) { g_create_error = "unknown game"; return SPB_ERR_ARG; }
  if (cfg->num_games == 0 || cfg->num_games > (1u << 22)) { g_create_error = "num_games out of range"; return SPB_ERR_ARG; }
  if (cfg->leaves_per_tree < 1 || cfg->leaves_per_tree > 16) { g_create_error = "leaves_per_tree must be in 1..16"; return SPB_ERR_ARG; }
  if (cfg->evaluator < SPB_EVAL_NET || cfg->evaluator > SPB_EVAL_UNIFORM) { g_create_error = "unknown evaluator"; return SPB_ERR_ARG; }
  if (!(cfg->c == cfg->c)) { g_create_error = "c is NaN"; return SPB_ERR_ARG; }
  spb_engine* e = new (std::nothrow) spb_engine();
  if (!e) { g_create_error = "out of host memory"; return SPB_ERR_NOMEM; }
  e->cfg = *cfg;
  if (e->cfg.max_nodes_per_tree == 0) e->cfg.max_nodes_per_tree = 16384;
  if (e->cfg.max_nodes_per_tree < 16 || e->cfg.max_nodes_per_tree > MAX_CAP) { g_create_error = "max_nodes_per_tree out of range"; delete e; return SPB_ERR_ARG; }
  int32_t rc = e->init();
  if (rc != SPB_OK) {
    g_create_error = e->err;
    e->destroy();
    delete e;
    return rc;
  }
  *out = e;
  return SPB_OK;
}

int32_t spb_destroy(spb_engine* e) {
  if (!e) return SPB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  e->destroy();
  delete e;
  return SPB_OK;
}

const char* spb_last_error(const spb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int32_t spb_load_weights(spb_engine* e, const void* blob, size_t n) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, blob && n > 8, "null / empty weight blob");
  HostNet net;
  std::string err;
  if (!parse_safetensors_net(blob, n, e->cfg.game, &net, &err)) { e->set_error("spb_load_weights: " + err); return SPB_ERR_WEIGHTS; }
  cudaStreamSynchronize(e->stream);
  if (!e->evaluator.upload(net, &err)) { e->set_error("spb_load_weights: " + err); return SPB_ERR_CUDA; }
  // The captured step graph holds the address of the previous weight image in its kernel nodes (hot swap between
  // generations, learner_concurrent.rs:158-159): drop it, the next search captures the new one.
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  e->last_eval_parity = -1;
  return SPB_OK;
}

int32_t spb_check_weights(int32_t game, const void* blob, size_t n, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (game != SPB_GAME_CONNECT4 && game != SPB_GAME_TICTACTOE) return SPB_ERR_ARG;
  if (!blob) return SPB_ERR_ARG;
  HostNet net;
  std::string msg;
  if (!parse_safetensors_net(blob, n, game, &net, &msg)) {
    if (err && err_cap) { std::strncpy(err, msg.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return SPB_ERR_WEIGHTS;
  }
  return SPB_OK;
}

int32_t spb_reset_games(spb_engine* e, const uint32_t* slots, uint32_t n, const spb_state* roots) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, n <= e->T.G, "more slots than games");
  if (n == 0) return SPB_OK;
  if (slots) for (uint32_t i = 0; i < n; ++i) ARG_CHECK(e, slots[i] < e->T.G, "slot out of range");
  size_t off_roots = ((size_t)n * 4 + 15) & ~(size_t)15;
  size_t bytes = off_roots + (size_t)n * sizeof(PState);
  int32_t rc = e->ensure_stage(bytes);
  if (rc) return rc;
  {
    auto* hs = static_cast<uint8_t*>(e->h_stage);
    if (slots) std::memcpy(hs, slots, (size_t)n * 4);
    if (roots) {
      PState* hp = reinterpret_cast<PState*>(hs + off_roots);
      for (uint32_t i = 0; i < n; ++i) hp[i] = ps_from_abi(roots[i]);
    }
  }
  {
    cudaError_t ce = cudaMemcpyAsync(e->d_stage, e->h_stage, bytes, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
    auto* ds = static_cast<uint8_t*>(e->d_stage);
    k_reset<<<(n + 127) / 128, 128, 0, e->stream>>>(e->T, e->P.hist_len, e->P.parked, slots ? reinterpret_cast<uint32_t*>(ds) : nullptr,
                                                    roots ? reinterpret_cast<PState*>(ds + off_roots) : nullptr, n);
    ++e->launches;
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);   // staging buffer is reused by the next call
    if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  }
  return SPB_OK;
}

int32_t spb_search(spb_engine* e, uint32_t num_searches) {
  ENGINE_GUARD(e);
  return DISPATCH_GAME(e, e->search_t<Connect4>(num_searches), e->search_t<TicTacToe>(num_searches));
}

static int32_t fetch_root_children(spb_engine* e) {
  k_root_children<<<(e->T.G + 127) / 128, 128, 0, e->stream>>>(e->T, e->d_rc_actions, e->d_rc_counts, e->d_rc_ids, e->d_rc_n);
  ++e->launches;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  return SPB_OK;
}

int32_t spb_root_children_all(spb_engine* e, uint8_t* actions, uint32_t* visit_counts, uint32_t* child_ids, uint32_t* n_children) {
  ENGINE_GUARD(e);
  int32_t rc = fetch_root_children(e);
  if (rc) return rc;
  const size_t G = e->T.G;
  // one pinned staging area, one sync
  size_t o_counts = 0, o_ids = G * SPB_MAX_ACTIONS * 4, o_n = o_ids + G * SPB_MAX_ACTIONS * 4, o_act = o_n + G * 4;
  size_t bytes = o_act + G * SPB_MAX_ACTIONS;
  rc = e->ensure_stage(bytes);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  cudaError_t ce = cudaSuccess;
  if (visit_counts) ce = cudaMemcpyAsync(hs + o_counts, e->d_rc_counts, G * SPB_MAX_ACTIONS * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && child_ids) ce = cudaMemcpyAsync(hs + o_ids, e->d_rc_ids, G * SPB_MAX_ACTIONS * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && n_children) ce = cudaMemcpyAsync(hs + o_n, e->d_rc_n, G * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && actions) ce = cudaMemcpyAsync(hs + o_act, e->d_rc_actions, G * SPB_MAX_ACTIONS, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (visit_counts) std::memcpy(visit_counts, hs + o_counts, G * SPB_MAX_ACTIONS * 4);
  if (child_ids) std::memcpy(child_ids, hs + o_ids, G * SPB_MAX_ACTIONS * 4);
  if (n_children) std::memcpy(n_children, hs + o_n, G * 4);
  if (actions) std::memcpy(actions, hs + o_act, G * SPB_MAX_ACTIONS);
  return SPB_OK;
}

int32_t spb_root_children(spb_engine* e, uint32_t slot, uint8_t* actions, uint32_t* visit_counts, uint32_t* child_ids, uint32_t* n_children) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, slot < e->T.G, "slot out of range");
  int32_t rc = fetch_root_children(e);
  if (rc) return rc;
  uint8_t a[SPB_MAX_ACTIONS]; uint32_t c[SPB_MAX_ACTIONS], ids[SPB_MAX_ACTIONS], n = 0;
  {
    cudaError_t ce = cudaMemcpyAsync(a, e->d_rc_actions + (size_t)slot * SPB_MAX_ACTIONS, SPB_MAX_ACTIONS, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(c, e->d_rc_counts + (size_t)slot * SPB_MAX_ACTIONS, SPB_MAX_ACTIONS * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(ids, e->d_rc_ids + (size_t)slot * SPB_MAX_ACTIONS, SPB_MAX_ACTIONS * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&n, e->d_rc_n + slot, 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  }
  for (uint32_t i = 0; i < n; ++i) {
    if (actions) actions[i] = a[i];
    if (visit_counts) visit_counts[i] = c[i];
    if (child_ids) child_ids[i] = ids[i];
  }
  if (n_children) *n_children = n;
  return SPB_OK;
}

int32_t spb_root_policy(spb_engine* e, uint32_t slot, float* policy) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, policy, "null policy");
  uint8_t a[SPB_MAX_ACTIONS]; uint32_t c[SPB_MAX_ACTIONS], n = 0;
  int32_t rc = spb_root_children(e, slot, a, c, nullptr, &n);
  if (rc) return rc;
  // mcts.rs:315-328: zero policy, set_prob(action, count as f32), normalize (ndarray sum order; counts are
  // integers < 2^24 so every order gives the same f32 sum), f32 divide.
  float p[SPB_MAX_ACTIONS] = {0};
  for (uint32_t i = 0; i < n; ++i) p[a[i]] = (float)c[i];
  float s = 0.0f;
  for (int i = 0; i < e->A; ++i) s += p[i];
  for (int i = 0; i < e->A; ++i) policy[i] = p[i] / s;
  return SPB_OK;
}

int32_t spb_advance(spb_engine* e, const uint32_t* slots, const uint32_t* node_ids, uint32_t n, spb_state* out_states) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, node_ids, "null node_ids");
  ARG_CHECK(e, n <= e->T.G, "more slots than games");
  if (n == 0) return SPB_OK;
  // validate against arena lengths (one D2H of n_nodes)
  std::vector<uint32_t> nn(e->T.G);
  {
    cudaError_t ce = cudaMemcpyAsync(nn.data(), e->T.n_nodes, (size_t)e->T.G * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  }
  std::vector<uint8_t> seen(e->T.G, 0);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t g = slots ? slots[i] : i;
    ARG_CHECK(e, g < e->T.G, "slot out of range");
    ARG_CHECK(e, !seen[g], "slot listed twice");
    seen[g] = 1;
    ARG_CHECK(e, node_ids[i] < nn[g], "node id out of range");
  }
  size_t o_ids = ((size_t)n * 4 + 15) & ~(size_t)15, o_out = o_ids * 2;
  size_t bytes = o_out + (size_t)n * sizeof(PState);
  int32_t rc = e->ensure_stage(bytes);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  auto* ds = static_cast<uint8_t*>(e->d_stage);
  if (slots) std::memcpy(hs, slots, (size_t)n * 4);
  std::memcpy(hs + o_ids, node_ids, (size_t)n * 4);
  cudaError_t ce = cudaMemcpyAsync(ds, hs, o_out, cudaMemcpyHostToDevice, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  const uint32_t blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  const uint32_t* dsl = slots ? reinterpret_cast<uint32_t*>(ds) : nullptr;
  const uint32_t* dn = reinterpret_cast<uint32_t*>(ds + o_ids);
  PState* dout = reinterpret_cast<PState*>(ds + o_out);
  if (e->cfg.game == SPB_GAME_CONNECT4) k_advance<Connect4><<<blocks, THREADS, 0, e->stream>>>(e->T, dsl, dn, n, dout);
  else k_advance<TicTacToe><<<blocks, THREADS, 0, e->stream>>>(e->T, dsl, dn, n, dout);
  ++e->launches;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hs + o_out, dout, (size_t)n * sizeof(PState), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (out_states) {
    const PState* hp = reinterpret_cast<const PState*>(hs + o_out);
    for (uint32_t i = 0; i < n; ++i) out_states[i] = ps_to_abi(hp[i]);
  }
  return SPB_OK;
}

int32_t spb_arena_len(spb_engine* e, uint32_t slot, uint32_t* out) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, slot < e->T.G && out, "bad argument");
  cudaError_t ce = cudaMemcpyAsync(out, e->T.n_nodes + slot, 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  return SPB_OK;
}

int32_t spb_get_state(spb_engine* e, uint32_t slot, uint32_t node_id, spb_state* out) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, slot < e->T.G && out, "bad argument");
  uint32_t len = 0;
  int32_t rc = spb_arena_len(e, slot, &len);
  if (rc) return rc;
  ARG_CHECK(e, node_id < len, "node id out of range");
  rc = e->ensure_stage(sizeof(PState));
  if (rc) return rc;
  if (e->cfg.game == SPB_GAME_CONNECT4) k_get_state<Connect4><<<1, 1, 0, e->stream>>>(e->T, slot, node_id, static_cast<PState*>(e->d_stage));
  else k_get_state<TicTacToe><<<1, 1, 0, e->stream>>>(e->T, slot, node_id, static_cast<PState*>(e->d_stage));
  ++e->launches;
  PState ps;
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&ps, e->d_stage, sizeof ps, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  *out = ps_to_abi(ps);
  return SPB_OK;
}

int32_t spb_node_stats(spb_engine* e, uint32_t slot, uint32_t node_id, uint32_t* visit_count, float* value_sum, float* prior,
                       uint32_t* first_child, uint32_t* n_children) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, slot < e->T.G, "slot out of range");
  uint32_t len = 0;
  int32_t rc = spb_arena_len(e, slot, &len);
  if (rc) return rc;
  ARG_CHECK(e, node_id < len, "node id out of range");
  rc = e->ensure_stage(sizeof(NodeRec));
  if (rc) return rc;
  k_node_stats<<<1, 1, 0, e->stream>>>(e->T, slot, node_id, static_cast<NodeRec*>(e->d_stage));
  ++e->launches;
  NodeRec r;
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&r, e->d_stage, sizeof r, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (visit_count) *visit_count = r.N;
  if (value_sum) *value_sum = r.W;
  if (prior) *prior = r.P;
  if (first_child) *first_child = info_fc(r.info);
  if (n_children) *n_children = info_nc(r.info);
  return SPB_OK;
}

}  // extern "C"

// ---- predict ------------------------------------------------------------------------------------
template <class G>
static int32_t predict_t(spb_engine* e, const spb_state* states, uint32_t n, float* policies, float* values, float* raw_logits) {
  const size_t o_states = 0, o_cnt = (size_t)n * sizeof(PState), o_out = (o_cnt + 16 + 15) & ~(size_t)15;
  const size_t o_pol = o_out + (size_t)n * G::EVAL_STRIDE * 4, o_val = o_pol + (size_t)n * G::A * 4;
  const size_t o_log = (o_val + (size_t)n * 4 + 15) & ~(size_t)15;
  const size_t bytes = o_log + (size_t)n * G::A * 4;
  int32_t rc = e->ensure_stage(bytes);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  auto* ds = static_cast<uint8_t*>(e->d_stage);
  PState* hp = reinterpret_cast<PState*>(hs + o_states);
  for (uint32_t i = 0; i < n; ++i) hp[i] = ps_from_abi(states[i]);
  *reinterpret_cast<uint32_t*>(hs + o_cnt) = n;
  cudaError_t ce = cudaMemcpyAsync(ds, hs, o_cnt + 4, cudaMemcpyHostToDevice, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  PState* dstates = reinterpret_cast<PState*>(ds + o_states);
  uint32_t* dcnt = reinterpret_cast<uint32_t*>(ds + o_cnt);
  float* dout = reinterpret_cast<float*>(ds + o_out);
  float* dlog = reinterpret_cast<float*>(ds + o_log);
  if (e->cfg.evaluator == SPB_EVAL_NET) {
    if (!e->evaluator.loaded()) { e->set_error("no weights loaded: call spb_load_weights first"); return SPB_ERR_STATE; }
    ce = cudaMemsetAsync(dlog, 0, (size_t)n * G::A * 4, e->stream);
    if (ce == cudaSuccess)
      ce = e->evaluator.launch(dstates, nullptr, dcnt, n, dout, G::EVAL_STRIDE, raw_logits ? dlog : nullptr,
                               (e->cfg.flags & SPB_FLAG_EVAL_SIMT) != 0, e->stream);
  } else {
    // identity work list: k_eval_builtin indexes states by list entry, so build 0..n-1 in the logits area
    std::vector<uint32_t> idl(n);
    for (uint32_t i = 0; i < n; ++i) idl[i] = i;
    ce = cudaMemcpyAsync(dlog, idl.data(), (size_t)n * 4, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) {
      if (e->cfg.evaluator == SPB_EVAL_DET)
        k_eval_builtin<G, SPB_EVAL_DET><<<(n + 255) / 256, 256, 0, e->stream>>>(dstates, reinterpret_cast<uint32_t*>(dlog), dcnt, dout);
      else
        k_eval_builtin<G, SPB_EVAL_UNIFORM><<<(n + 255) / 256, 256, 0, e->stream>>>(dstates, reinterpret_cast<uint32_t*>(dlog), dcnt, dout);
      ce = cudaGetLastError();
    }
  }
  ++e->launches;
  if (ce != cudaSuccess) { e->set_error(std::string("evaluator: ") + cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  k_mask_policies<G><<<(n + 127) / 128, 128, 0, e->stream>>>(dstates, n, dout, reinterpret_cast<float*>(ds + o_pol), reinterpret_cast<float*>(ds + o_val));
  ++e->launches;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hs + o_pol, ds + o_pol, bytes - o_pol, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(std::string("predict: ") + cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (policies) std::memcpy(policies, hs + o_pol, (size_t)n * G::A * 4);
  if (values) std::memcpy(values, hs + o_val, (size_t)n * 4);
  if (raw_logits && e->cfg.evaluator == SPB_EVAL_NET) std::memcpy(raw_logits, hs + o_log, (size_t)n * G::A * 4);
  return SPB_OK;
}

extern "C" {

int32_t spb_predict(spb_engine* e, const spb_state* states, uint32_t n, float* policies, float* values, float* raw_logits) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, states || n == 0, "null states");
  if (n == 0) return SPB_OK;
  return DISPATCH_GAME(e, predict_t<Connect4>(e, states, n, policies, values, raw_logits),
                       predict_t<TicTacToe>(e, states, n, policies, values, raw_logits));
}

// ---- game rules ---------------------------------------------------------------------------------
static int32_t upload_states(spb_engine* e, const spb_state* states, uint32_t n, size_t extra_bytes) {
  int32_t rc = e->ensure_stage((size_t)n * sizeof(PState) + extra_bytes + 64);
  if (rc) return rc;
  PState* hp = static_cast<PState*>(e->h_stage);
  for (uint32_t i = 0; i < n; ++i) hp[i] = ps_from_abi(states[i]);
  return SPB_OK;
}

int32_t spb_game_next_states(spb_engine* e, const spb_state* states, const uint8_t* actions, uint32_t n, spb_state* out_states, int32_t* err) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, states && actions && out_states && err, "null argument");
  if (n == 0) return SPB_OK;
  const size_t o_act = (size_t)n * sizeof(PState), o_out = (o_act + n + 15) & ~(size_t)15, o_err = o_out + (size_t)n * sizeof(PState);
  int32_t rc = upload_states(e, states, n, o_err + (size_t)n * 4);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  auto* ds = static_cast<uint8_t*>(e->d_stage);
  std::memcpy(hs + o_act, actions, n);
  cudaError_t ce = cudaMemcpyAsync(ds, hs, o_out, cudaMemcpyHostToDevice, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (e->cfg.game == SPB_GAME_CONNECT4)
    k_game_next<Connect4><<<(n + 127) / 128, 128, 0, e->stream>>>((PState*)ds, ds + o_act, n, (PState*)(ds + o_out), (int32_t*)(ds + o_err));
  else
    k_game_next<TicTacToe><<<(n + 127) / 128, 128, 0, e->stream>>>((PState*)ds, ds + o_act, n, (PState*)(ds + o_out), (int32_t*)(ds + o_err));
  ++e->launches;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hs + o_out, ds + o_out, (size_t)n * (sizeof(PState) + 4), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  const PState* hp = reinterpret_cast<const PState*>(hs + o_out);
  const int32_t* he = reinterpret_cast<const int32_t*>(hs + o_err);
  for (uint32_t i = 0; i < n; ++i) { out_states[i] = ps_to_abi(hp[i]); err[i] = he[i]; }
  return SPB_OK;
}

int32_t spb_game_valid_actions(spb_engine* e, const spb_state* states, uint32_t n, uint32_t* masks) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, states && masks, "null argument");
  if (n == 0) return SPB_OK;
  const size_t o_out = (size_t)n * sizeof(PState);
  int32_t rc = upload_states(e, states, n, o_out + (size_t)n * 4);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  auto* ds = static_cast<uint8_t*>(e->d_stage);
  cudaError_t ce = cudaMemcpyAsync(ds, hs, o_out, cudaMemcpyHostToDevice, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  if (e->cfg.game == SPB_GAME_CONNECT4) k_game_valid<Connect4><<<(n + 127) / 128, 128, 0, e->stream>>>((PState*)ds, n, (uint32_t*)(ds + o_out));
  else k_game_valid<TicTacToe><<<(n + 127) / 128, 128, 0, e->stream>>>((PState*)ds, n, (uint32_t*)(ds + o_out));
  ++e->launches;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(masks, ds + o_out, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  return SPB_OK;
}

int32_t spb_game_encode(spb_engine* e, const spb_state* states, uint32_t n, float* out) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, states && out, "null argument");
  if (n == 0) return SPB_OK;
  const bool c4 = e->cfg.game == SPB_GAME_CONNECT4;
  const size_t E = c4 ? 3 * 6 * 7 : 27;
  const size_t o_out = (size_t)n * sizeof(PState);
  int32_t rc = upload_states(e, states, n, o_out + (size_t)n * E * 4);
  if (rc) return rc;
  auto* hs = static_cast<uint8_t*>(e->h_stage);
  auto* ds = static_cast<uint8_t*>(e->d_stage);
  cudaError_t ce = cudaMemcpyAsync(ds, hs, o_out, cudaMemcpyHostToDevice, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  const size_t total = (size_t)n * E;
  if (c4) k_game_encode<Connect4><<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>((PState*)ds, n, (float*)(ds + o_out));
  else k_game_encode<TicTacToe><<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>((PState*)ds, n, (float*)(ds + o_out));
  ++e->launches;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, ds + o_out, total * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  return SPB_OK;
}

// ---- self-play ----------------------------------------------------------------------------------
int32_t spb_selfplay_step(spb_engine* e, int32_t rule, float temperature, uint64_t seed, const spb_state* restart_roots, uint32_t* n_finished) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, rule == SPB_MOVE_GREEDY_LAST_MAX || rule == SPB_MOVE_TEMPERATURE, "unknown move rule");
  const uint32_t G = e->T.G;
  PState* droots = nullptr;
  if (restart_roots) {
    int32_t rc = upload_states(e, restart_roots, G, 0);
    if (rc) return rc;
    cudaError_t ce = cudaMemcpyAsync(e->d_stage, e->h_stage, (size_t)G * sizeof(PState), cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
    droots = static_cast<PState*>(e->d_stage);
  }
  cudaError_t ce = cudaMemsetAsync(e->P.finished, 0, 4, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  const uint32_t blocks = (G + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  if (e->cfg.game == SPB_GAME_CONNECT4) k_selfplay_step<Connect4><<<blocks, THREADS, 0, e->stream>>>(e->T, e->P, rule, temperature, seed, droots);
  else k_selfplay_step<TicTacToe><<<blocks, THREADS, 0, e->stream>>>(e->T, e->P, rule, temperature, seed, droots);
  ++e->launches;
  uint32_t fin = 0;
  ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&fin, e->P.finished, 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  int32_t rc = e->check_device_errors();
  if (n_finished) *n_finished = fin;
  return rc;
}

int32_t spb_drain_trajectories(spb_engine* e, spb_position* buf, size_t capacity, size_t* written, uint64_t* game_ids) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, written, "null written");
  unsigned long long cur = 0;
  cudaError_t ce = cudaMemcpyAsync(&cur, e->P.out_cursor, 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  size_t n = (size_t)std::min<unsigned long long>(cur, e->P.out_cap);
  *written = n;
  if (!buf) return SPB_OK;                        // size query
  ARG_CHECK(e, capacity >= n, "trajectory buffer too small");
  if (n == 0) return SPB_OK;
  std::vector<spb_position> pos(n);
  std::vector<unsigned long long> ids(n);
  ce = cudaMemcpyAsync(pos.data(), e->P.out, n * sizeof(spb_position), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ids.data(), e->P.out_game, n * 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(e->P.out_cursor, 0, 8, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  // deterministic order: game id, then ply (games finish in a nondeterministic order on the device)
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    if (ids[a] != ids[b]) return ids[a] < ids[b];
    return pos[a].ply < pos[b].ply;
  });
  for (size_t i = 0; i < n; ++i) {
    buf[i] = pos[order[i]];
    if (game_ids) game_ids[i] = ids[order[i]];
  }
  return SPB_OK;
}

// ---- counters -----------------------------------------------------------------------------------
int32_t spb_get_counters(spb_engine* e, spb_counters* out) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, out, "null out");
  std::memset(out, 0, sizeof *out);
  unsigned long long c[CTR_COUNT], live = 0;
  cudaError_t ce = cudaMemsetAsync(e->d_misc, 0, 8, e->stream);
  if (ce == cudaSuccess) {
    k_nodes_live<<<(e->T.G + 127) / 128, 128, 0, e->stream>>>(e->T, e->d_misc);
    ++e->launches;
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(c, e->T.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&live, e->d_misc, 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  out->simulations = c[CTR_SIMS];
  out->evaluations = c[CTR_EVALS];
  out->terminal_leaves = c[CTR_TERMINAL];
  out->path_length_sum = c[CTR_PATHSUM];
  out->children_created = c[CTR_CHILDREN];
  out->nodes_live = live;
  out->kernel_launches = e->launches;
  return SPB_OK;
}

int32_t spb_reset_counters(spb_engine* e) {
  ENGINE_GUARD(e);
  cudaError_t ce = cudaMemsetAsync(e->T.counters, 0, CTR_COUNT * 8, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  e->launches = 0;
  return SPB_OK;
}

int32_t spb_last_search_timing(spb_engine* e, float* search_ms, float* evaluator_ms, uint32_t* evaluator_launches) {
  if (!e) return SPB_ERR_ARG;
  if (search_ms) *search_ms = e->last_search_ms;
  if (evaluator_ms) *evaluator_ms = e->last_eval_ms;
  if (evaluator_launches) *evaluator_launches = e->last_eval_launches;
  return SPB_OK;
}

int32_t spb_time_evaluator(spb_engine* e, uint32_t iters, float* avg_ms, uint32_t* n_positions, double* flops_per_position) {
  ENGINE_GUARD(e);
  ARG_CHECK(e, iters > 0 && avg_ms, "bad argument");
  if (e->cfg.evaluator != SPB_EVAL_NET || !e->evaluator.loaded()) { e->set_error("needs the network evaluator with weights loaded"); return SPB_ERR_STATE; }
  const uint32_t slots = e->T.G * e->T.K;
  const uint32_t* cnt = &e->T.eval_count[2];
  const bool simt = (e->cfg.flags & SPB_FLAG_EVAL_SIMT) != 0;
  cudaError_t ce = cudaSuccess;
  // Lock-step pipeline: the work list of the last simulation step.  Asynchronous pipeline (no step lists): the latest
  // evaluated leaf of every slot, as one static list of G positions.
  const uint32_t* list = e->T.eval_list;
  if (e->last_eval_parity < 0) {
    list = nullptr;
    ce = cudaMemcpyAsync(&e->T.eval_count[2], &slots, 4, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  }
  for (int w = 0; w < 2 && ce == cudaSuccess; ++w)    // warm-up
    ce = e->evaluator.launch(e->T.leaf_state, list, cnt, slots, e->T.eval_out, e->eval_stride, nullptr, simt, e->stream, /*overlap=*/false);
  if (ce == cudaSuccess) ce = cudaEventRecord(e->ev0, e->stream);
  for (uint32_t i = 0; i < iters && ce == cudaSuccess; ++i)
    ce = e->evaluator.launch(e->T.leaf_state, list, cnt, slots, e->T.eval_out, e->eval_stride, nullptr, simt, e->stream, /*overlap=*/false);
  if (ce == cudaSuccess) ce = cudaEventRecord(e->ev1, e->stream);
  uint32_t n = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&n, cnt, 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  float ms = 0.f;
  if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  if (ce != cudaSuccess) { e->set_error(std::string("spb_time_evaluator: ") + cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  e->launches += iters + 2;
  *avg_ms = ms / (float)iters;
  if (n_positions) *n_positions = n;
  if (flops_per_position) *flops_per_position = e->evaluator.flops_per_position();
  return SPB_OK;
}

int32_t spb_last_async_stats(spb_engine* e, uint64_t* out, uint32_t n) {
  if (!e || !out) return SPB_ERR_ARG;
  for (uint32_t i = 0; i < n; ++i) out[i] = i < (uint32_t)ASTAT_COUNT ? e->last_async_stats[i] : 0;
  return SPB_OK;
}

int32_t spb_synchronize(spb_engine* e) {
  ENGINE_GUARD(e);
  cudaError_t ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { e->set_error(cudaGetErrorString(ce)); return SPB_ERR_CUDA; }
  return SPB_OK;
}

}  // extern "C"
