// chess_engine.cu — chess trees and search behind the C ABI (include/selfplay_b200.h, spb_chess_*): BASELINE config 5.
// Device primitives: chess_tree.cuh (select / expand / backup / warp move generation), chess.cuh (rules).
//   Mcts::search (ref: src/mcts.rs:196-332)      -> k_chess_search_fused (DetEval / uniform: all simulations of a tree in
//                                                   one warp) and, for the network, the lock-step pair k_chess_select ->
//                                                   [chess_net.cu] -> k_chess_finish: exactly the reference's loop, one
//                                                   evaluator batch per simulation step over all trees
//   Tree::use_subtree (:161-192)                 -> k_chess_advance (breadth-first compaction into the other arena)
//   Tree::with_root_state (:86-89)               -> k_chess_reset
#include "chess_engine.hpp"
#include "chess_net.cuh"

namespace spb {
namespace chess {

constexpr int WARPS = 4;
constexpr int THREADS = WARPS * 32;
constexpr uint32_t SPLIT_MIN_TREES = 128;   // the network pipeline runs as two half-loops on two streams from 2 x this many trees on

__global__ void k_chess_reset(CTrees T, const uint32_t* slots, const Pos* roots, const unsigned long long* hist, uint32_t n) {
  const uint32_t i = blockIdx.x;
  if (i >= n) return;
  const uint32_t g = slots ? slots[i] : i;
  Pos root = roots ? roots[i] : start_position();
  if (!hist) root.hist_len = 0;
  if (root.hist_len > SPB_CHESS_MAX_HISTORY) root.hist_len = SPB_CHESS_MAX_HISTORY;
  if (hist)
    for (uint32_t k = threadIdx.x; k < root.hist_len; k += blockDim.x)
      T.root_hist[(size_t)g * SPB_CHESS_MAX_HISTORY + k] = hist[(size_t)i * SPB_CHESS_MAX_HISTORY + k];
  if (threadIdx.x == 0) {
    T.root[g] = root;
    T.buf[g] = 0;
    T.live[g] = 1;
    T.n_nodes[g] = 1;
    T.leaf_depth[g] = 0;
    T.rec[0][(size_t)g * T.cap] = make_uint4(0u, 0u, 0u, 0u);
    T.meta[0][(size_t)g * T.cap] = make_uint2(NO_PARENT, make_meta(MOVE_NONE, 0, ST_UNVISITED));
    T.hash[0][(size_t)g * T.cap] = 0ull;
  }
}

__device__ __forceinline__ void flush_counters(const CTrees& T, const unsigned long long* ctr, int lane) {
  if (lane == 0)
    for (int i = 0; i < CTR_COUNT; ++i)
      if (ctr[i]) atomicAdd(&T.counters[i], ctr[i]);
}

// What a simulation finds at its leaf (get_value_and_terminated, chess.rs:168-174, evaluated once per node): the legal
// moves in ws.moves, their hash, the repetition count; returns the node's status (ST_EXPANDED = not terminal).
__device__ __forceinline__ uint32_t visit_leaf(const CTrees& T, uint32_t g, const Pos& pos, int depth, int lane, WarpScratch& ws, int& n,
                                               unsigned long long& h, uint32_t& reps) {
  n = warp_legal_moves(pos, lane, ws, T.error);
  h = warp_list_hash(ws, n, lane);
  reps = warp_repetitions(h, T.root_hist + (size_t)g * SPB_CHESS_MAX_HISTORY, T.root[g].hist_len, ws, depth, lane);
  if (n == 0) return in_check(pos, pos.side) ? ST_WON : ST_TIED;     // chess.rs:156-157
  if (reps >= 3 || pos.fifty >= 100) return ST_TIED;                  // :159-160
  return ST_EXPANDED;
}

// Fused search (DetEval / uniform): all simulations of one tree inside one warp.
template <int EVAL>
__global__ void __launch_bounds__(THREADS) k_chess_search_fused(CTrees T, uint32_t num_searches) {
  __shared__ WarpScratch s_ws[WARPS];
  const int lane = threadIdx.x & 31;
  WarpScratch& ws = s_ws[threadIdx.x >> 5];
  const uint32_t g = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (g >= T.G || !T.live[g]) return;
  const uint32_t b = T.buf[g];
  uint4* rec = T.rec[b] + (size_t)g * T.cap;
  uint2* meta = T.meta[b] + (size_t)g * T.cap;
  unsigned long long* hash = T.hash[b] + (size_t)g * T.cap;
  const Pos root = T.root[g];
  uint32_t n_nodes = T.n_nodes[g];
  unsigned long long ctr[CTR_COUNT] = {0, 0, 0, 0, 0};
  for (uint32_t s = 0; s < num_searches; ++s) {                        // mcts.rs:214
    Pos pos = root;
    uint32_t node, lmeta;
    int depth;
    if (!descend(rec, meta, hash, T.c, lane, ws, pos, node, depth, lmeta, T.error)) break;
    ctr[CTR_SIMS] += 1;
    ctr[CTR_PATHSUM] += (unsigned)depth;
    uint32_t st = meta_status(lmeta);
    int n = 0;
    unsigned long long h = 0;
    uint32_t reps = 0;
    if (st == ST_UNVISITED) {
      st = visit_leaf(T, g, pos, depth, lane, ws, n, h, reps);
      if (st != ST_EXPANDED && lane == 0) {                            // terminal: remember it (mcts.rs:245 on every later visit)
        meta[node].y = make_meta(meta_move(lmeta), 0, st);
        hash[node] = h;
      }
    }
    float v;
    if (st != ST_EXPANDED) {
      v = st == ST_WON ? 1.0f : 0.0f;                                  // chess.rs:172
      ctr[CTR_TERMINAL] += 1;
    } else {
      // predict (model/mod.rs:36-98) by a built-in evaluator + mask_invalid_actions (chess.rs:251-271)
      const uint64_t dh = det_hash(pos);
      float sum = 0.0f;
      if (EVAL == SPB_EVAL_DET) {
        for (int i = lane; i < n; i += 32) sum += det_raw_prob(dh, policy_index(pos.side, ws.moves[i]));   // dyadic: exact in any order
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        v = det_value(dh);
      } else {
        sum = (float)n;
        v = 0.0f;
      }
      ctr[CTR_EVALS] += 1;
      const int side = pos.side;
      const bool ok = expand(rec, meta, hash, T.cap, n_nodes, node, ws, n, h, lane, [&](int i) {
        const float raw = EVAL == SPB_EVAL_DET ? det_raw_prob(dh, policy_index(side, ws.moves[i])) : 1.0f;
        return __fdiv_rn(raw, sum);
      });
      if (!ok) {
        if (lane == 0) atomicOr(T.error, (uint32_t)ERRBIT_POOL);
        break;
      }
      ctr[CTR_CHILDREN] += (unsigned)n;
    }
    __syncwarp();
    backup(rec, ws, depth, v, lane);
    __syncwarp();
  }
  if (lane == 0) T.n_nodes[g] = n_nodes;
  flush_counters(T, ctr, lane);
}

// ---- lock-step pipeline for the network --------------------------------------------------------------------------
// k_chess_select: the select of one simulation per tree; terminal leaves are backed up at once, the others are stored
// with their legal moves as the tree's pending leaf and appended to the evaluator's work list (mcts.rs:236-252).
__global__ void __launch_bounds__(THREADS) k_chess_select(CTrees T, TreeRange R) {
  __shared__ WarpScratch s_ws[WARPS];
  const int lane = threadIdx.x & 31;
  WarpScratch& ws = s_ws[threadIdx.x >> 5];
  const uint32_t g = R.g0 + blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (g >= R.g1 || !T.live[g]) return;
  const uint32_t b = T.buf[g];
  uint4* rec = T.rec[b] + (size_t)g * T.cap;
  uint2* meta = T.meta[b] + (size_t)g * T.cap;
  unsigned long long* hash = T.hash[b] + (size_t)g * T.cap;
  unsigned long long ctr[CTR_COUNT] = {0, 0, 0, 0, 0};
  Pos pos = T.root[g];
  uint32_t node, lmeta;
  int depth;
  if (!descend(rec, meta, hash, T.c, lane, ws, pos, node, depth, lmeta, T.error)) return;
  ctr[CTR_SIMS] += 1;
  ctr[CTR_PATHSUM] += (unsigned)depth;
  uint32_t st = meta_status(lmeta);
  int n = 0;
  unsigned long long h = 0;
  uint32_t reps = 0;
  if (st == ST_UNVISITED) {
    st = visit_leaf(T, g, pos, depth, lane, ws, n, h, reps);
    if (st != ST_EXPANDED && lane == 0) {
      meta[node].y = make_meta(meta_move(lmeta), 0, st);
      hash[node] = h;
    }
  }
  if (st != ST_EXPANDED) {
    ctr[CTR_TERMINAL] += 1;
    __syncwarp();
    backup(rec, ws, depth, st == ST_WON ? 1.0f : 0.0f, lane);
  } else {
    ctr[CTR_EVALS] += 1;
    for (int i = lane; i < n; i += 32) T.leaf_moves[(size_t)g * MAX_MOVES + i] = ws.moves[i];
    for (int d = lane; d <= depth; d += 32) T.path[(size_t)g * MAX_DEPTH + d] = ws.path[d];
    if (lane == 0) {
      T.leaf_pos[g] = pos;
      T.leaf_node[g] = node;
      T.leaf_depth[g] = (uint32_t)depth | LEAF_PENDING;
      T.leaf_nmoves[g] = (uint32_t)n;
      T.leaf_reps[g] = reps;
      T.leaf_hash[g] = h;
      R.list[atomicAdd(R.count, 1u)] = g;
    }
  }
  flush_counters(T, ctr, lane);
}

// k_chess_finish: expand + backup of the evaluated leaf (mcts.rs:268-284).  Model::predict's tail (model/mod.rs:64-95):
// softmax over all 4,672 logits, then mask_invalid_actions = keep the legal cells and divide by their sum
// (chess.rs:251-271).  The normaliser of the softmax cancels in that division, so only the legal logits are read:
// prior_i = exp(l_i - m) / sum_legal exp(l_j - m), m = max over the legal logits.
// RAW = true (parity harness, SPB_FLAG_FORCE_SPLIT): eval_logits holds the raw probabilities of a built-in evaluator for the
// legal cells (k_chess_builtin_eval) and the prior is raw / sum, exactly as in the fused kernel.
template <bool RAW>
__global__ void __launch_bounds__(THREADS) k_chess_finish(CTrees T, TreeRange R) {
  __shared__ WarpScratch s_ws[WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  WarpScratch& ws = s_ws[w];
  // the exponentials of the legal logits live in the scratch's path-hash array, which this kernel does not use: 10 KB of
  // shared memory per block, so that a block fits next to a resident k_conv CTA (chess_net.cu)
  static_assert(sizeof(ws.phash) >= MAX_MOVES * sizeof(float), "phash too small for the priors");
  float* s_ev = reinterpret_cast<float*>(ws.phash);
  const uint32_t g = R.g0 + blockIdx.x * WARPS + w;
  if (g >= R.g1 || !T.live[g]) return;
  const uint32_t ld = T.leaf_depth[g];
  if (!(ld & LEAF_PENDING)) return;
  const int depth = (int)(ld & 0xFFu);
  const uint32_t b = T.buf[g];
  uint4* rec = T.rec[b] + (size_t)g * T.cap;
  uint2* meta = T.meta[b] + (size_t)g * T.cap;
  unsigned long long* hash = T.hash[b] + (size_t)g * T.cap;
  const uint32_t node = T.leaf_node[g];
  const int n = (int)T.leaf_nmoves[g];
  const int side = T.leaf_pos[g].side;
  const float* logits = T.eval_logits + (size_t)g * LOGIT_STRIDE;
  for (int i = lane; i < n; i += 32) ws.moves[i] = T.leaf_moves[(size_t)g * MAX_MOVES + i];
  for (int d = lane; d <= depth; d += 32) ws.path[d] = T.path[(size_t)g * MAX_DEPTH + d];
  __syncwarp();
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) {
    const float l = logits[policy_index(side, ws.moves[i])];
    s_ev[i] = l;
    mx = fmaxf(mx, l);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
  for (int i = lane; i < n; i += 32) {
    const float e = RAW ? s_ev[i] : __expf(s_ev[i] - mx);
    s_ev[i] = e;
    sum += e;                                                        // RAW: dyadic values, exact in any order
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  uint32_t n_nodes = T.n_nodes[g];
  const float* ev = s_ev;
  const bool ok = expand(rec, meta, hash, T.cap, n_nodes, node, ws, n, T.leaf_hash[g], lane, [&](int i) { return __fdiv_rn(ev[i], sum); });
  if (!ok) {
    if (lane == 0) { atomicOr(T.error, (uint32_t)ERRBIT_POOL); T.leaf_depth[g] = 0; }
    return;
  }
  if (lane == 0) {
    T.n_nodes[g] = n_nodes;
    T.leaf_depth[g] = 0;
    atomicAdd(&T.counters[CTR_CHILDREN], (unsigned long long)n);
  }
  __syncwarp();
  backup(rec, ws, depth, T.eval_value[g], lane);
}

// Built-in evaluators for the lock-step pipeline (parity harness): the raw probabilities of the legal cells and the value of
// every pending leaf, where the network would have written its logits.
template <int EVAL>
__global__ void __launch_bounds__(THREADS) k_chess_builtin_eval(CTrees T, TreeRange R) {
  const int lane = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (i >= *R.count) return;
  const uint32_t g = R.list[i];
  const Pos pos = T.leaf_pos[g];
  const uint64_t dh = det_hash(pos);
  const int n = (int)T.leaf_nmoves[g];
  float* logits = T.eval_logits + (size_t)g * LOGIT_STRIDE;
  for (int k = lane; k < n; k += 32) {
    const int idx = policy_index(pos.side, T.leaf_moves[(size_t)g * MAX_MOVES + k]);
    logits[idx] = EVAL == SPB_EVAL_DET ? det_raw_prob(dh, idx) : 1.0f;
  }
  if (lane == 0) T.eval_value[g] = EVAL == SPB_EVAL_DET ? det_value(dh) : 0.0f;
}

// ---- results / re-rooting ----------------------------------------------------------------------------------------
__global__ void k_chess_root_children(CTrees T, uint16_t* moves, uint32_t* counts, uint32_t* ids, uint32_t* n_out) {
  const uint32_t g = blockIdx.x;
  if (g >= T.G) return;
  uint32_t nc = 0, fc = 0;
  const uint32_t b = T.buf[g];
  const uint4* rec = T.rec[b] + (size_t)g * T.cap;
  const uint2* meta = T.meta[b] + (size_t)g * T.cap;
  if (T.live[g] && meta_status(meta[0].y) == ST_EXPANDED) { nc = meta_nc(meta[0].y); fc = rec[0].w; }
  for (uint32_t i = threadIdx.x; i < MAX_MOVES; i += blockDim.x) {
    const bool in = i < nc;
    moves[(size_t)g * MAX_MOVES + i] = in ? (uint16_t)meta_move(meta[fc + i].y) : MOVE_NONE;
    counts[(size_t)g * MAX_MOVES + i] = in ? rec[fc + i].x : 0u;
    ids[(size_t)g * MAX_MOVES + i] = in ? fc + i : 0u;
  }
  if (threadIdx.x == 0) n_out[g] = nc;
}

// use_subtree (mcts.rs:161-192) for a child of the root: breadth-first copy of the kept subtree into the other arena,
// ids assigned in queue order with every node's children contiguous and in their old order.  One warp per tree; a
// window of 32 copied nodes per round: a scan of their child counts gives each its new first_child, then the warp
// copies the children of the 32 nodes one node after the other (lanes stride over up to 218 children).  A copied node
// keeps its OLD first_child in rec.w until its own turn in the window.  The root position advances by the child's move
// and the hash of the old root's move list joins the game history (chess.rs:121-122).
__global__ void __launch_bounds__(THREADS) k_chess_advance(CTrees T, const uint32_t* slots, const uint32_t* node_ids, uint32_t n, Pos* out_states,
                                                           int32_t* out_err) {
  const int lane = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint32_t g = slots ? slots[i] : i;
  const uint32_t b = T.buf[g], nb = b ^ 1u;
  const uint4* orec = T.rec[b] + (size_t)g * T.cap;
  const uint2* ometa = T.meta[b] + (size_t)g * T.cap;
  const unsigned long long* ohash = T.hash[b] + (size_t)g * T.cap;
  uint4* nrec = T.rec[nb] + (size_t)g * T.cap;
  uint2* nmeta = T.meta[nb] + (size_t)g * T.cap;
  unsigned long long* nhash = T.hash[nb] + (size_t)g * T.cap;
  const uint32_t id = node_ids[i];
  Pos root = T.root[g];
  int32_t err = SPB_OK;
  if (!T.live[g] || id == 0 || id >= T.n_nodes[g] || ometa[id].x != 0u) err = SPB_ERR_ARG;        // must be a child of the root
  else if (root.hist_len >= SPB_CHESS_MAX_HISTORY) err = SPB_ERR_STATE;
  if (err != SPB_OK) {
    if (lane == 0) { out_err[i] = err; if (out_states) out_states[i] = root; }
    return;
  }
  const Move mv = (Move)meta_move(ometa[id].y);
  if (lane == 0) {
    T.root_hist[(size_t)g * SPB_CHESS_MAX_HISTORY + root.hist_len] = ohash[0];
    const uint32_t hl = root.hist_len + 1;
    root = apply_move(root, mv);
    root.hist_len = hl;
    T.root[g] = root;
    if (out_states) out_states[i] = root;
    out_err[i] = SPB_OK;
    nrec[0] = orec[id];
    nmeta[0] = make_uint2(NO_PARENT, ometa[id].y);
    nhash[0] = ohash[id];
  }
  __syncwarp();
  uint32_t n_new = 1, j0 = 0;
  while (j0 < n_new) {
    const uint32_t in_window = min(32u, n_new - j0);                   // nodes copied so far that still await their turn
    const uint32_t j = j0 + lane;
    uint32_t nc = 0, old_fc = 0;
    if ((uint32_t)lane < in_window) {
      const uint32_t my = nmeta[j].y;
      if (meta_status(my) == ST_EXPANDED) { nc = meta_nc(my); old_fc = nrec[j].w; }
    }
    int total;
    const uint32_t new_fc = n_new + (uint32_t)warp_excl_scan((int)nc, lane, &total);
    if (n_new + (uint32_t)total > T.cap) {                             // cannot happen: the subtree is part of an arena of <= cap nodes
      if (lane == 0) atomicOr(T.error, (uint32_t)ERRBIT_BOUNDS | (43u << 8));
      break;
    }
    if (nc) nrec[j].w = new_fc;
    for (uint32_t l = 0; l < in_window; ++l) {
      const uint32_t cnc = __shfl_sync(0xffffffffu, nc, l), cold = __shfl_sync(0xffffffffu, old_fc, l), cnew = __shfl_sync(0xffffffffu, new_fc, l);
      for (uint32_t k = lane; k < cnc; k += 32) {
        nrec[cnew + k] = orec[cold + k];
        nmeta[cnew + k] = make_uint2(j0 + l, ometa[cold + k].y);
        nhash[cnew + k] = ohash[cold + k];
      }
    }
    n_new += (uint32_t)total;
    j0 += in_window;
    __syncwarp();
  }
  if (lane == 0) {
    T.n_nodes[g] = n_new;
    T.buf[g] = (uint8_t)nb;
    T.leaf_depth[g] = 0;
  }
}

// arena[node_id].state (mcts.rs:22): the moves from the root to the node, replayed.
__global__ void k_chess_get_state(CTrees T, uint32_t g, uint32_t node_id, Pos* out, int32_t* err) {
  const uint32_t b = T.buf[g];
  const uint2* meta = T.meta[b] + (size_t)g * T.cap;
  if (!T.live[g] || node_id >= T.n_nodes[g]) { *err = SPB_ERR_ARG; return; }
  Move mv[MAX_DEPTH];
  int d = 0;
  for (uint32_t v = node_id; v != 0; v = meta[v].x) {
    if (d >= MAX_DEPTH) { *err = SPB_ERR_STATE; return; }
    mv[d++] = (Move)meta_move(meta[v].y);
  }
  Pos p = T.root[g];
  const uint32_t hl = p.hist_len;
  while (d > 0) p = apply_move(p, mv[--d]);
  p.hist_len = hl;   // the history that travels with the root; the path's own positions are not exported
  *out = p;
  *err = SPB_OK;
}

__global__ void k_chess_node_stats(CTrees T, uint32_t g, uint32_t node_id, uint32_t* out) {
  const uint32_t b = T.buf[g];
  if (!T.live[g] || node_id >= T.n_nodes[g]) { out[7] = 1; return; }
  const uint4 r = T.rec[b][(size_t)g * T.cap + node_id];
  const uint2 m = T.meta[b][(size_t)g * T.cap + node_id];
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
  out[3] = meta_status(m.y) == ST_EXPANDED ? r.w : 0u;
  out[4] = meta_status(m.y) == ST_EXPANDED ? meta_nc(m.y) : 0u;
  out[5] = meta_move(m.y);
  out[6] = meta_status(m.y);
  out[7] = 0;
}

__global__ void k_chess_nodes_live(CTrees T, unsigned long long* out) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < T.G && T.live[g]) { atomicAdd(out, (unsigned long long)T.n_nodes[g]); atomicMax(out + 1, (unsigned long long)T.n_nodes[g]); }
}

}  // namespace chess
}  // namespace spb

using namespace spb;
namespace ch = spb::chess;

int32_t spb_chess_engine::check_device_errors() {
  uint32_t bits = 0;
  CH_CUDA(this, cudaMemcpyAsync(&bits, T.error, 4, cudaMemcpyDeviceToHost, stream));
  CH_CUDA(this, cudaStreamSynchronize(stream));
  if (!bits) return SPB_OK;
  CH_CUDA(this, cudaMemsetAsync(T.error, 0, 4, stream));
  if (bits & ERRBIT_POOL) { set_error("a chess tree outgrew max_nodes_per_tree"); return SPB_ERR_POOL; }
  set_error("chess kernel check failed, site " + std::to_string((bits >> 8) & 0xFF));
  return SPB_ERR_STATE;
}

extern "C" {

int32_t spb_chess_create(const spb_config* cfg, spb_chess_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return SPB_ERR_ARG; }
  *out = nullptr;
  if (cfg->abi_version != SPB_ABI_VERSION) { g_create_error = "abi_version mismatch"; return SPB_ERR_ARG; }
  if (cfg->game != SPB_GAME_CHESS) { g_create_error = "spb_chess_create needs game = SPB_GAME_CHESS"; return SPB_ERR_ARG; }
  if (cfg->num_games == 0 || cfg->num_games > (1u << 20)) { g_create_error = "num_games out of range"; return SPB_ERR_ARG; }
  if (cfg->leaves_per_tree != 1) { g_create_error = "chess search runs one leaf per tree per step"; return SPB_ERR_ARG; }
  if (cfg->evaluator < SPB_EVAL_NET || cfg->evaluator > SPB_EVAL_UNIFORM) { g_create_error = "unknown evaluator"; return SPB_ERR_ARG; }
  if (!(cfg->c == cfg->c)) { g_create_error = "c is NaN"; return SPB_ERR_ARG; }
  spb_chess_engine* e = new (std::nothrow) spb_chess_engine();
  if (!e) { g_create_error = "out of host memory"; return SPB_ERR_NOMEM; }
  e->cfg = *cfg;
  if (e->cfg.max_nodes_per_tree == 0) e->cfg.max_nodes_per_tree = 32768;
  auto fail = [&](int32_t rc, const std::string& msg) {
    g_create_error = msg;
    spb_chess_destroy(e);
    return rc;
  };
  if (e->cfg.max_nodes_per_tree < 512 || e->cfg.max_nodes_per_tree > (1u << 24)) return fail(SPB_ERR_ARG, "max_nodes_per_tree out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SPB_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SPB_ERR_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(SPB_ERR_CUDA, "cudaSetDevice failed");
  if (prop.major != 10) return fail(SPB_ERR_CUDA, "device is not sm_100 (B200); this library is built for sm_100a only");
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&e->ev0) != cudaSuccess ||
      cudaEventCreate(&e->ev1) != cudaSuccess)
    return fail(SPB_ERR_CUDA, "stream / event creation failed");
  ch::CTrees& T = e->T;
  T.G = cfg->num_games; T.cap = e->cfg.max_nodes_per_tree; T.c = cfg->c;
  const size_t G = T.G, pool = G * (size_t)T.cap;
  int32_t rc = SPB_OK;
  for (int b = 0; b < 2 && !rc; ++b) {
    if (!rc) rc = e->dalloc(&T.rec[b], pool);
    if (!rc) rc = e->dalloc(&T.meta[b], pool);
    if (!rc) rc = e->dalloc(&T.hash[b], pool);
  }
  if (!rc) rc = e->dalloc(&T.root, G);
  if (!rc) rc = e->dalloc(&T.root_hist, G * SPB_CHESS_MAX_HISTORY);
  if (!rc) rc = e->dalloc(&T.n_nodes, G);
  if (!rc) rc = e->dalloc(&T.buf, G);
  if (!rc) rc = e->dalloc(&T.live, G);
  if (!rc) rc = e->dalloc(&T.counters, (size_t)CTR_COUNT);
  if (!rc) rc = e->dalloc(&T.error, 1);
  if (!rc) rc = e->dalloc(&T.leaf_node, G);
  if (!rc) rc = e->dalloc(&T.leaf_depth, G);
  if (!rc) rc = e->dalloc(&T.path, G * ch::MAX_DEPTH);
  if (!rc) rc = e->dalloc(&T.leaf_pos, G);
  if (!rc) rc = e->dalloc(&T.leaf_moves, G * ch::MAX_MOVES);
  if (!rc) rc = e->dalloc(&T.leaf_nmoves, G);
  if (!rc) rc = e->dalloc(&T.leaf_reps, G);
  if (!rc) rc = e->dalloc(&T.leaf_hash, G);
  if (!rc) rc = e->dalloc(&T.eval_list, G);
  if (!rc) rc = e->dalloc(&T.eval_count, 2);
  if (!rc) rc = e->dalloc(&T.eval_value, G);
  if (!rc && (cfg->evaluator == SPB_EVAL_NET || (cfg->flags & SPB_FLAG_FORCE_SPLIT))) rc = e->dalloc(&T.eval_logits, G * (size_t)ch::LOGIT_STRIDE);
  if (!rc) rc = e->dalloc(&e->d_rc_moves, G * ch::MAX_MOVES);
  if (!rc) rc = e->dalloc(&e->d_rc_counts, G * ch::MAX_MOVES);
  if (!rc) rc = e->dalloc(&e->d_rc_ids, G * ch::MAX_MOVES);
  if (!rc) rc = e->dalloc(&e->d_rc_n, G);
  if (!rc) rc = e->dalloc(&e->d_misc, 4);
  if (!rc) rc = e->dalloc(&e->d_reset_slots, G);
  if (!rc) rc = e->dalloc(&e->d_reset_roots, G);
  if (!rc) rc = e->dalloc(&e->d_reset_hist, G * SPB_CHESS_MAX_HISTORY);
  if (!rc) rc = e->dalloc(&e->d_adv_ids, G);
  if (!rc) rc = e->dalloc(&e->d_adv_err, G);
  if (rc) return fail(rc, e->err);
  cudaMemsetAsync(T.live, 0, G, e->stream);
  cudaMemsetAsync(T.buf, 0, G, e->stream);
  cudaMemsetAsync(T.n_nodes, 0, G * 4, e->stream);
  cudaMemsetAsync(T.leaf_depth, 0, G * 4, e->stream);
  cudaMemsetAsync(T.counters, 0, CTR_COUNT * 8, e->stream);
  cudaMemsetAsync(T.error, 0, 4, e->stream);
  cudaMemsetAsync(T.eval_count, 0, 8, e->stream);
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return fail(SPB_ERR_CUDA, "device initialisation failed");
  *out = e;
  return SPB_OK;
}

int32_t spb_chess_destroy(spb_chess_engine* e) {
  if (!e) return SPB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  ch::net_destroy(e->net);
  for (void* p : e->allocs) cudaFree(p);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return SPB_OK;
}

const char* spb_chess_last_error(const spb_chess_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int32_t spb_chess_reset_games(spb_chess_engine* e, const uint32_t* slots, uint32_t n, const spb_chess_state* roots, const uint64_t* history) {
  CH_GUARD(e);
  CH_ARG(e, n <= e->T.G, "more slots than games");
  CH_ARG(e, !(history && !roots), "history without roots");
  if (n == 0) return SPB_OK;
  if (slots) for (uint32_t i = 0; i < n; ++i) CH_ARG(e, slots[i] < e->T.G, "slot out of range");
  if (roots) for (uint32_t i = 0; i < n; ++i) {
    CH_ARG(e, roots[i].side <= 1 && roots[i].ep <= 64, "malformed root state");
    uint64_t kings[2] = {roots[i].piece[5] & roots[i].color[0], roots[i].piece[5] & roots[i].color[1]};
    CH_ARG(e, kings[0] && !(kings[0] & (kings[0] - 1)) && kings[1] && !(kings[1] & (kings[1] - 1)), "root state needs one king per side");
    CH_ARG(e, history || roots[i].hist_len == 0, "root with hist_len > 0 needs its history");
    CH_ARG(e, roots[i].hist_len <= SPB_CHESS_MAX_HISTORY, "hist_len > SPB_CHESS_MAX_HISTORY");
  }
  uint32_t* d_slots = slots ? e->d_reset_slots : nullptr;
  ch::Pos* d_roots = roots ? e->d_reset_roots : nullptr;
  unsigned long long* d_hist = history ? e->d_reset_hist : nullptr;
  if (slots) CH_CUDA(e, cudaMemcpyAsync(d_slots, slots, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  if (roots) CH_CUDA(e, cudaMemcpyAsync(d_roots, roots, (size_t)n * sizeof(ch::Pos), cudaMemcpyHostToDevice, e->stream));
  if (history) CH_CUDA(e, cudaMemcpyAsync(d_hist, history, (size_t)n * SPB_CHESS_MAX_HISTORY * 8, cudaMemcpyHostToDevice, e->stream));
  ch::k_chess_reset<<<n, 64, 0, e->stream>>>(e->T, d_slots, d_roots, d_hist, n);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  return SPB_OK;
}

int32_t spb_chess_search(spb_chess_engine* e, uint32_t num_searches) {
  CH_GUARD(e);
  if (num_searches == 0) return SPB_OK;
  const uint32_t blocks = (e->T.G + ch::WARPS - 1) / ch::WARPS;
  CH_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  const ch::TreeRange all{0u, e->T.G, e->T.eval_list, e->T.eval_count};
  if (e->cfg.evaluator == SPB_EVAL_NET) {
    CH_ARG(e, e->net && ch::net_loaded(e->net), "no weights loaded (spb_chess_load_weights)");
    // mcts.rs:214: one evaluator batch per simulation step.  Trees never interact, so the lock-step loop runs separately
    // over the two halves of the trees, on two streams: while the convolutions of one half hold the tensor pipes, the
    // select / expand + backup kernels of the other half run in the SMs' spare issue slots and shared memory (a k_conv CTA
    // leaves room for them).  Results are those of one loop over all trees (tests: node for node against the oracle).
    const uint32_t G = e->T.G, G0 = (G + 1u) / 2u;
    const bool split = G >= 2u * ch::SPLIT_MIN_TREES && !(e->cfg.flags & SPB_FLAG_LOCKSTEP);
    const ch::TreeRange half[2] = {{0u, split ? G0 : G, e->T.eval_list, e->T.eval_count}, {G0, G, e->T.eval_list + G0, e->T.eval_count + 1}};
    const int nh = split ? 2 : 1;
    cudaStream_t st[2] = {e->stream, e->stream2};
    if (split) {
      CH_CUDA(e, cudaEventRecord(e->ev_fork, e->stream));
      CH_CUDA(e, cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
    }
    auto enqueue = [&]() -> int32_t {
      for (uint32_t s = 0; s < num_searches; ++s) {
        for (int h = 0; h < nh; ++h) {
          const ch::TreeRange& R = half[h];
          const uint32_t hb = (R.g1 - R.g0 + ch::WARPS - 1) / ch::WARPS;
          CH_CUDA(e, cudaMemsetAsync(R.count, 0, 4, st[h]));
          ch::k_chess_select<<<hb, ch::THREADS, 0, st[h]>>>(e->T, R);
          CH_CUDA(e, cudaGetLastError());
          uint32_t launched = 0;
          const int32_t rc = ch::net_forward_leaves(e, h, R.list, R.count, st[h], &launched);
          if (rc) return rc;
          ch::k_chess_finish<false><<<hb, ch::THREADS, 0, st[h]>>>(e->T, R);
          CH_CUDA(e, cudaGetLastError());
          e->launches += 2 + launched;
        }
      }
      return SPB_OK;
    };
    const int32_t rc_loop = enqueue();
    if (split) {                                                       // the second stream rejoins the engine's stream, also after an error
      cudaError_t je = cudaEventRecord(e->ev_join, e->stream2);
      if (je == cudaSuccess) je = cudaStreamWaitEvent(e->stream, e->ev_join, 0);
      if (je != cudaSuccess && rc_loop == SPB_OK) { e->set_error(std::string("chess search: ") + cudaGetErrorString(je)); return SPB_ERR_CUDA; }
    }
    if (rc_loop) return rc_loop;
  } else if (e->cfg.flags & SPB_FLAG_FORCE_SPLIT) {
    // parity harness: the built-in evaluators through the kernels of the network pipeline (select -> evaluate -> finish)
    for (uint32_t s = 0; s < num_searches; ++s) {
      CH_CUDA(e, cudaMemsetAsync(e->T.eval_count, 0, 4, e->stream));
      ch::k_chess_select<<<blocks, ch::THREADS, 0, e->stream>>>(e->T, all);
      if (e->cfg.evaluator == SPB_EVAL_DET) ch::k_chess_builtin_eval<SPB_EVAL_DET><<<blocks, ch::THREADS, 0, e->stream>>>(e->T, all);
      else ch::k_chess_builtin_eval<SPB_EVAL_UNIFORM><<<blocks, ch::THREADS, 0, e->stream>>>(e->T, all);
      ch::k_chess_finish<true><<<blocks, ch::THREADS, 0, e->stream>>>(e->T, all);
      CH_CUDA(e, cudaGetLastError());
      e->launches += 3;
    }
  } else {
    if (e->cfg.evaluator == SPB_EVAL_DET) ch::k_chess_search_fused<SPB_EVAL_DET><<<blocks, ch::THREADS, 0, e->stream>>>(e->T, num_searches);
    else ch::k_chess_search_fused<SPB_EVAL_UNIFORM><<<blocks, ch::THREADS, 0, e->stream>>>(e->T, num_searches);
    CH_CUDA(e, cudaGetLastError());
    ++e->launches;
  }
  CH_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  const int32_t rc = e->check_device_errors();
  if (rc) return rc;
  CH_CUDA(e, cudaEventElapsedTime(&e->last_search_ms, e->ev0, e->ev1));
  return SPB_OK;
}

int32_t spb_chess_last_search_ms(spb_chess_engine* e, float* ms) {
  CH_GUARD(e);
  CH_ARG(e, ms, "null argument");
  *ms = e->last_search_ms;
  return SPB_OK;
}

static int32_t fetch_root_children(spb_chess_engine* e) {
  ch::k_chess_root_children<<<e->T.G, 64, 0, e->stream>>>(e->T, e->d_rc_moves, e->d_rc_counts, e->d_rc_ids, e->d_rc_n);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  return SPB_OK;
}

int32_t spb_chess_root_children_all(spb_chess_engine* e, uint16_t* moves, uint32_t* visit_counts, uint32_t* child_ids, uint32_t* n_children) {
  CH_GUARD(e);
  int32_t rc = fetch_root_children(e);
  if (rc) return rc;
  const size_t G = e->T.G, M = ch::MAX_MOVES;
  if (moves) CH_CUDA(e, cudaMemcpyAsync(moves, e->d_rc_moves, G * M * 2, cudaMemcpyDeviceToHost, e->stream));
  if (visit_counts) CH_CUDA(e, cudaMemcpyAsync(visit_counts, e->d_rc_counts, G * M * 4, cudaMemcpyDeviceToHost, e->stream));
  if (child_ids) CH_CUDA(e, cudaMemcpyAsync(child_ids, e->d_rc_ids, G * M * 4, cudaMemcpyDeviceToHost, e->stream));
  if (n_children) CH_CUDA(e, cudaMemcpyAsync(n_children, e->d_rc_n, G * 4, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  return SPB_OK;
}

int32_t spb_chess_root_children(spb_chess_engine* e, uint32_t slot, uint16_t* moves, uint32_t* visit_counts, uint32_t* child_ids, uint32_t* n_children) {
  CH_GUARD(e);
  CH_ARG(e, slot < e->T.G, "slot out of range");
  CH_ARG(e, n_children, "null argument");
  int32_t rc = fetch_root_children(e);
  if (rc) return rc;
  const size_t M = ch::MAX_MOVES;
  if (moves) CH_CUDA(e, cudaMemcpyAsync(moves, e->d_rc_moves + slot * M, M * 2, cudaMemcpyDeviceToHost, e->stream));
  if (visit_counts) CH_CUDA(e, cudaMemcpyAsync(visit_counts, e->d_rc_counts + slot * M, M * 4, cudaMemcpyDeviceToHost, e->stream));
  if (child_ids) CH_CUDA(e, cudaMemcpyAsync(child_ids, e->d_rc_ids + slot * M, M * 4, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaMemcpyAsync(n_children, e->d_rc_n + slot, 4, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  return SPB_OK;
}

// Policy of Mcts::search's result (mcts.rs:315-328): the root children's visit counts scattered into the 73x8x8 table by
// set_prob (chess.rs:505-514), then normalize (divide by the sum, chess.rs:516-518).
int32_t spb_chess_root_policy(spb_chess_engine* e, uint32_t slot, float* out) {
  CH_GUARD(e);
  CH_ARG(e, out, "null argument");
  uint16_t mv[ch::MAX_MOVES];
  uint32_t cnt[ch::MAX_MOVES], n = 0;
  int32_t rc = spb_chess_root_children(e, slot, mv, cnt, nullptr, &n);
  if (rc) return rc;
  ch::Pos root;
  CH_CUDA(e, cudaMemcpyAsync(&root, e->T.root + slot, sizeof root, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  std::memset(out, 0, sizeof(float) * SPB_CHESS_POLICY_SIZE);
  float sum = 0.0f;
  for (uint32_t i = 0; i < n; ++i) sum += (float)cnt[i];               // integers: exact in any order
  for (uint32_t i = 0; i < n; ++i) out[ch::policy_index(root.side, mv[i])] = (float)cnt[i] / sum;
  return SPB_OK;
}

int32_t spb_chess_advance(spb_chess_engine* e, const uint32_t* slots, const uint32_t* child_ids, uint32_t n, spb_chess_state* out_states) {
  CH_GUARD(e);
  CH_ARG(e, child_ids && n <= e->T.G, "bad argument");
  if (n == 0) return SPB_OK;
  if (slots) for (uint32_t i = 0; i < n; ++i) CH_ARG(e, slots[i] < e->T.G, "slot out of range");
  // persistent device staging (every call ends with a stream synchronisation): the per-move call of the self-play loop
  uint32_t* d_slots = slots ? e->d_reset_slots : nullptr;
  uint32_t* d_ids = e->d_adv_ids;
  ch::Pos* d_out = e->d_reset_roots;
  int32_t* d_err = e->d_adv_err;
  if (slots) CH_CUDA(e, cudaMemcpyAsync(d_slots, slots, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  CH_CUDA(e, cudaMemcpyAsync(d_ids, child_ids, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  ch::k_chess_advance<<<(n + ch::WARPS - 1) / ch::WARPS, ch::THREADS, 0, e->stream>>>(e->T, d_slots, d_ids, n, d_out, d_err);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  std::vector<int32_t> err(n);
  CH_CUDA(e, cudaMemcpyAsync(err.data(), d_err, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
  if (out_states) CH_CUDA(e, cudaMemcpyAsync(out_states, d_out, (size_t)n * sizeof(ch::Pos), cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  for (uint32_t i = 0; i < n; ++i)
    if (err[i] != SPB_OK) {
      e->set_error(err[i] == SPB_ERR_STATE ? "game history full (SPB_CHESS_MAX_HISTORY plies)" : "spb_chess_advance: not a child of the root");
      return err[i];
    }
  return SPB_OK;
}

int32_t spb_chess_get_state(spb_chess_engine* e, uint32_t slot, uint32_t node_id, spb_chess_state* out) {
  CH_GUARD(e);
  CH_ARG(e, out && slot < e->T.G, "bad argument");
  ch::Scratch sc;
  ch::Pos* d_out = sc.alloc<ch::Pos>(1);
  int32_t* d_err = sc.alloc<int32_t>(1);
  if (!d_out || !d_err) { e->set_error("chess: out of device memory"); return SPB_ERR_NOMEM; }
  ch::k_chess_get_state<<<1, 1, 0, e->stream>>>(e->T, slot, node_id, d_out, d_err);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  int32_t err = 0;
  CH_CUDA(e, cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaMemcpyAsync(out, d_out, sizeof(ch::Pos), cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  if (err) { e->set_error("spb_chess_get_state: node out of range"); return err; }
  return SPB_OK;
}

int32_t spb_chess_arena_len(spb_chess_engine* e, uint32_t slot, uint32_t* out) {
  CH_GUARD(e);
  CH_ARG(e, out && slot < e->T.G, "bad argument");
  CH_CUDA(e, cudaMemcpyAsync(out, e->T.n_nodes + slot, 4, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  return SPB_OK;
}

int32_t spb_chess_node_stats(spb_chess_engine* e, uint32_t slot, uint32_t node_id, uint32_t* visit_count, float* value_sum, float* prior,
                             uint32_t* first_child, uint32_t* n_children, uint16_t* move, uint8_t* status) {
  CH_GUARD(e);
  CH_ARG(e, slot < e->T.G, "slot out of range");
  ch::Scratch sc;
  uint32_t* d = sc.alloc<uint32_t>(8);
  if (!d) { e->set_error("chess: out of device memory"); return SPB_ERR_NOMEM; }
  ch::k_chess_node_stats<<<1, 1, 0, e->stream>>>(e->T, slot, node_id, d);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  uint32_t h[8];
  CH_CUDA(e, cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  CH_ARG(e, h[7] == 0, "node out of range");
  if (visit_count) *visit_count = h[0];
  if (value_sum) std::memcpy(value_sum, &h[1], 4);
  if (prior) std::memcpy(prior, &h[2], 4);
  if (first_child) *first_child = h[3];
  if (n_children) *n_children = h[4];
  if (move) *move = (uint16_t)h[5];
  // status in the header's terms: a node that was never reached as a leaf has not been classified yet -> reported Ongoing
  if (status) *status = h[6] == ch::ST_TIED ? SPB_STATUS_TIED : (h[6] == ch::ST_WON ? SPB_STATUS_WON : SPB_STATUS_ONGOING);
  return SPB_OK;
}

int32_t spb_chess_get_counters(spb_chess_engine* e, spb_counters* out) {
  CH_GUARD(e);
  CH_ARG(e, out, "null argument");
  unsigned long long c[CTR_COUNT], live[2] = {0, 0};
  CH_CUDA(e, cudaMemsetAsync(e->d_misc, 0, 16, e->stream));
  ch::k_chess_nodes_live<<<(e->T.G + 127) / 128, 128, 0, e->stream>>>(e->T, e->d_misc);
  CH_CUDA(e, cudaGetLastError());
  ++e->launches;
  CH_CUDA(e, cudaMemcpyAsync(c, e->T.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaMemcpyAsync(live, e->d_misc, 16, cudaMemcpyDeviceToHost, e->stream));
  CH_CUDA(e, cudaStreamSynchronize(e->stream));
  std::memset(out, 0, sizeof *out);
  out->simulations = c[CTR_SIMS]; out->evaluations = c[CTR_EVALS]; out->terminal_leaves = c[CTR_TERMINAL];
  out->path_length_sum = c[CTR_PATHSUM]; out->children_created = c[CTR_CHILDREN];
  out->nodes_live = live[0]; out->kernel_launches = e->launches;
  out->reserved[0] = live[1];   // largest arena
  return SPB_OK;
}

int32_t spb_chess_reset_counters(spb_chess_engine* e) {
  CH_GUARD(e);
  CH_CUDA(e, cudaMemsetAsync(e->T.counters, 0, CTR_COUNT * 8, e->stream));
  e->launches = 0;
  return SPB_OK;
}

}  // extern "C"
