// games.cuh — bitboard game rules for the device (sm_100a).
//
// Replaces the `State` trait implementations of the reference:
//   src/game/connect_four.rs:127-283 and src/game/tictactoe.rs:127-241.
// A position is two 64-bit words (16 B, one LDG.128):
//   x : stones of Player::X                       (bits 0..47)
//   o : stones of Player::O                       (bits 0..47)
//       | num_actions_played << 48 (6 bits) | status << 56 (2 bits) | current_player << 60
// Connect4 bit = col*7 + row (row 0 = bottom, bit col*7+6 is always 0, so vertical and diagonal
// shifts never wrap between columns); tic-tac-toe bit = row*3 + col.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/selfplay_b200.h"

namespace spb {

constexpr uint64_t STONE_MASK = (1ull << 48) - 1;
constexpr int META_N_SHIFT = 48, META_STATUS_SHIFT = 56, META_PLAYER_SHIFT = 60;

struct __align__(16) PState {
  uint64_t x, o;
};

__host__ __device__ __forceinline__ uint32_t ps_status(const PState& s) { return (uint32_t)(s.o >> META_STATUS_SHIFT) & 3u; }
__host__ __device__ __forceinline__ uint32_t ps_player(const PState& s) { return (uint32_t)(s.o >> META_PLAYER_SHIFT) & 1u; }
__host__ __device__ __forceinline__ uint32_t ps_num_actions(const PState& s) { return (uint32_t)(s.o >> META_N_SHIFT) & 63u; }
__host__ __device__ __forceinline__ uint64_t ps_x(const PState& s) { return s.x & STONE_MASK; }
__host__ __device__ __forceinline__ uint64_t ps_o(const PState& s) { return s.o & STONE_MASK; }
__host__ __device__ __forceinline__ uint64_t ps_mine(const PState& s) { return ps_player(s) ? ps_o(s) : ps_x(s); }
__host__ __device__ __forceinline__ uint64_t ps_opp(const PState& s) { return ps_player(s) ? ps_x(s) : ps_o(s); }

__host__ __device__ __forceinline__ PState ps_make(uint64_t x, uint64_t o, uint32_t player, uint32_t n, uint32_t status) {
  PState s;
  s.x = x & STONE_MASK;
  s.o = (o & STONE_MASK) | ((uint64_t)(n & 63u) << META_N_SHIFT) | ((uint64_t)(status & 3u) << META_STATUS_SHIFT) |
        ((uint64_t)(player & 1u) << META_PLAYER_SHIFT);
  return s;
}
__host__ __device__ __forceinline__ PState ps_from_abi(const spb_state& a) {
  return ps_make(a.stones[0], a.stones[1], a.current_player, a.num_actions_played, a.status);
}
__host__ __device__ __forceinline__ spb_state ps_to_abi(const PState& s) {
  spb_state a;
  a.stones[0] = ps_x(s);
  a.stones[1] = ps_o(s);
  a.current_player = (uint8_t)ps_player(s);
  a.num_actions_played = (uint8_t)ps_num_actions(s);
  a.status = (uint8_t)ps_status(s);
  for (int i = 0; i < 5; ++i) a.reserved[i] = 0;
  return a;
}

// ---- Connect4 ------------------------------------------------------------------------------
struct Connect4 {
  static constexpr int GAME = SPB_GAME_CONNECT4;
  static constexpr int A = 7;          // actions (columns)
  static constexpr int ROWS = 6, COLS = 7;
  static constexpr int MAX_DEPTH = 44; // root + at most 42 plies below it (+1 spare)
  static constexpr int EVAL_STRIDE = 8;   // floats per evaluator output record: 7 probs + value

  // get_valid_actions, connect_four.rs:213-225: columns whose top cell (row 5) is empty, ascending.
  __device__ __forceinline__ static uint32_t valid_mask(const PState& s) {
    if (ps_status(s) != SPB_STATUS_ONGOING) return 0u;
    uint64_t occ = ps_x(s) | ps_o(s);
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 7; ++c) m |= (uint32_t)((~occ >> (c * 7 + 5)) & 1ull) << c;
    return m;
  }

  // One action's legality (same rule as valid_mask, for a position that is still being played): lane a of a warp
  // tests column a, a ballot then gives the legal mask without a 7-step loop.
  __device__ __forceinline__ static bool action_legal(const PState& s, int a) {
    return (((ps_x(s) | ps_o(s)) >> (a * 7 + 5)) & 1ull) == 0;
  }

  // get_winner, connect_four.rs:139-180, literally: any four in the latest ROW (either player),
  // any four in the latest COLUMN, and the (row+i, col+i) diagonal windows i in [start,end]
  // (:164-166).  The anti-diagonal is NOT checked — reference behaviour, reproduced on purpose.
  __device__ __forceinline__ static bool has_winner(uint64_t x, uint64_t o, int row, int col) {
    const uint64_t row_starts = (1ull << row) | (1ull << (7 + row)) | (1ull << (14 + row)) | (1ull << (21 + row));
    const uint64_t col_starts = 7ull << (col * 7);   // start rows 0..2
    int mn = min(col, row);
    int start = max(-4, -mn);
    int end = min(0, min(7 - (col + 4), 6 - (row + 4)));
    uint64_t diag_starts = 0;
    for (int i = start; i <= end; ++i) diag_starts |= 1ull << ((col + i) * 7 + (row + i));
    bool won = false;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uint64_t b = p ? o : x;
      uint64_t h = b & (b >> 7) & (b >> 14) & (b >> 21);
      uint64_t v = b & (b >> 1) & (b >> 2) & (b >> 3);
      uint64_t d = b & (b >> 8) & (b >> 16) & (b >> 24);
      won |= ((h & row_starts) | (v & col_starts) | (d & diag_starts)) != 0;
    }
    return won;
  }

  // get_next_state, connect_four.rs:190-211.  Returns false where the reference returns Err.
  __device__ __forceinline__ static bool next_state(const PState& s, int action, PState* out) {
    if (ps_status(s) != SPB_STATUS_ONGOING) return false;        // :209
    if (action < 0 || action >= 7) return false;
    uint64_t x = ps_x(s), o = ps_o(s);
    uint32_t colbits = (uint32_t)(((x | o) >> (action * 7)) & 0x3Full);
    if (colbits == 0x3Fu) return false;                          // :193 column already filled
    int row = __ffs((int)(~colbits & 0x3Fu)) - 1;                // :135 first empty row from the bottom
    uint32_t player = ps_player(s);
    uint64_t bit = 1ull << (action * 7 + row);
    if (player) o |= bit; else x |= bit;                         // :196
    uint32_t n = ps_num_actions(s) + 1;                          // :198
    uint32_t status = SPB_STATUS_ONGOING;
    if (has_winner(x, o, row, action)) status = SPB_STATUS_WON;  // :200
    else if (n == 42) status = SPB_STATUS_TIED;                  // :202
    *out = ps_make(x, o, player ^ 1u, n, status);                // :197
    return true;
  }

  // Stone placement only (status comes from the node record during descent).
  __device__ __forceinline__ static PState place(const PState& s, int action, uint32_t status) {
    uint64_t x = ps_x(s), o = ps_o(s);
    uint32_t colbits = (uint32_t)(((x | o) >> (action * 7)) & 0x3Full);
    int row = __ffs((int)(~colbits & 0x3Fu)) - 1;
    uint64_t bit = 1ull << (action * 7 + row);
    uint32_t player = ps_player(s);
    if (player) o |= bit; else x |= bit;
    return ps_make(x, o, player ^ 1u, ps_num_actions(s) + 1, status);
  }

  // get_encoding, connect_four.rs:242-259: out[plane][row][col].
  __device__ __forceinline__ static float encode_cell(const PState& s, int plane, int row, int col) {
    uint64_t mine = ps_mine(s), opp = ps_opp(s);
    int b = col * 7 + row;
    uint32_t m = (uint32_t)(mine >> b) & 1u, p = (uint32_t)(opp >> b) & 1u;
    uint32_t v = plane == 0 ? m : (plane == 1 ? p : (1u ^ m ^ p));
    return v ? 1.0f : 0.0f;
  }

  // ndarray sum() over 7 contiguous f32: plain left-to-right from 0.0 (connect_four.rs:276).
  __device__ __forceinline__ static float masked_sum(const float* m) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; ++i) acc = __fadd_rn(acc, m[i]);
    return acc;
  }
};

// ---- Tic-tac-toe ---------------------------------------------------------------------------
struct TicTacToe {
  static constexpr int GAME = SPB_GAME_TICTACTOE;
  static constexpr int A = 9;
  static constexpr int ROWS = 3, COLS = 3;
  static constexpr int MAX_DEPTH = 12;
  static constexpr int EVAL_STRIDE = 16;  // 9 probs + value, padded

  // get_valid_actions, tictactoe.rs:169-182 (row-major).
  __device__ __forceinline__ static uint32_t valid_mask(const PState& s) {
    if (ps_status(s) != SPB_STATUS_ONGOING) return 0u;
    return (uint32_t)(~(ps_x(s) | ps_o(s))) & 0x1FFu;
  }
  __device__ __forceinline__ static bool action_legal(const PState& s, int a) {
    return (((uint32_t)(ps_x(s) | ps_o(s)) >> a) & 1u) == 0;
  }

  // get_next_state, tictactoe.rs:135-167.
  __device__ __forceinline__ static bool next_state(const PState& s, int action, PState* out) {
    if (ps_status(s) != SPB_STATUS_ONGOING) return false;       // :165
    if (action < 0 || action >= 9) return false;
    uint64_t x = ps_x(s), o = ps_o(s);
    if (((x | o) >> action) & 1ull) return false;               // :138
    uint32_t player = ps_player(s);
    if (player) o |= 1ull << action; else x |= 1ull << action;
    uint32_t n = ps_num_actions(s) + 1;
    uint32_t m = (uint32_t)(player ? o : x);                    // the mover's stones
    int r = action / 3, c = action % 3;
    uint32_t rowl = 7u << (3 * r), coll = 0x49u << c;
    bool win = (m & rowl) == rowl || (m & coll) == coll;        // :146-150 (placed cell is Some => all the mover's)
    if (r == c) win |= (m & 0x111u) == 0x111u;                  // :152
    if ((r == 1 && c == 1) || abs(r - c) == 2) win |= (m & 0x54u) == 0x54u;   // :154
    uint32_t status = win ? SPB_STATUS_WON : (n == 9 ? SPB_STATUS_TIED : SPB_STATUS_ONGOING);   // :157-160
    *out = ps_make(x, o, player ^ 1u, n, status);
    return true;
  }

  __device__ __forceinline__ static PState place(const PState& s, int action, uint32_t status) {
    uint64_t x = ps_x(s), o = ps_o(s);
    uint32_t player = ps_player(s);
    if (player) o |= 1ull << action; else x |= 1ull << action;
    return ps_make(x, o, player ^ 1u, ps_num_actions(s) + 1, status);
  }

  __device__ __forceinline__ static float encode_cell(const PState& s, int plane, int row, int col) {
    uint64_t mine = ps_mine(s), opp = ps_opp(s);
    int b = row * 3 + col;
    uint32_t m = (uint32_t)(mine >> b) & 1u, p = (uint32_t)(opp >> b) & 1u;
    uint32_t v = plane == 0 ? m : (plane == 1 ? p : (1u ^ m ^ p));
    return v ? 1.0f : 0.0f;
  }

  // ndarray sum() over 9 contiguous f32 (tictactoe.rs:233): eight stride-8 partial sums combined as
  // ((((0+(x0+x4))+(x1+x5))+(x2+x6))+(x3+x7)), then the tail x8.
  __device__ __forceinline__ static float masked_sum(const float* m) {
    float acc = 0.0f;
    acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(0.0f, m[0]), __fadd_rn(0.0f, m[4])));
    acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(0.0f, m[1]), __fadd_rn(0.0f, m[5])));
    acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(0.0f, m[2]), __fadd_rn(0.0f, m[6])));
    acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(0.0f, m[3]), __fadd_rn(0.0f, m[7])));
    acc = __fadd_rn(acc, m[8]);
    return acc;
  }
};

// get_value_and_terminated, connect_four.rs:231-240 / tictactoe.rs:188-197.
__device__ __forceinline__ float terminal_value(uint32_t status) { return status == SPB_STATUS_WON ? -1.0f : 0.0f; }

// mask_invalid_actions, connect_four.rs:261-279 / tictactoe.rs:218-236: p*mask / sum(p*mask).
template <class G>
__device__ __forceinline__ void mask_renorm(uint32_t legal, const float* probs, float* out) {
  float m[G::A];
#pragma unroll
  for (int a = 0; a < G::A; ++a) m[a] = __fmul_rn(probs[a], (legal >> a & 1u) ? 1.0f : 0.0f);
  float s = G::masked_sum(m);
#pragma unroll
  for (int a = 0; a < G::A; ++a) out[a] = __fdiv_rn(m[a], s);
}

// The same, for ONE action: expand needs only the prior of the child a lane creates (one IEEE division instead of A).
template <class G>
__device__ __forceinline__ float mask_renorm_one(uint32_t legal, const float* probs, int action) {
  float m[G::A];
#pragma unroll
  for (int a = 0; a < G::A; ++a) m[a] = __fmul_rn(probs[a], (legal >> a & 1u) ? 1.0f : 0.0f);
  const float s = G::masked_sum(m);
  float mine = 0.0f;
#pragma unroll
  for (int a = 0; a < G::A; ++a) if (a == action) mine = m[a];
  return __fdiv_rn(mine, s);
}

// ---- DetEval ---------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class G>
__device__ __forceinline__ void det_eval(const PState& s, float* probs, float* value) {
  uint64_t h = splitmix64(ps_mine(s) ^ splitmix64(ps_opp(s)));
#pragma unroll
  for (int a = 0; a < G::A; ++a) probs[a] = __fdiv_rn((float)(1u + (uint32_t)((h >> (4 * a)) & 7ull)), 64.0f);
  *value = __fdiv_rn(__fsub_rn((float)((h >> 40) & 0xFFull), 128.0f), 128.0f);
}
template <class G>
__device__ __forceinline__ void uniform_eval(const PState&, float* probs, float* value) {
#pragma unroll
  for (int a = 0; a < G::A; ++a) probs[a] = 1.0f;
  *value = 0.0f;
}

// i-th set bit (0-based) of a small mask.
__device__ __forceinline__ int nth_set_bit(uint32_t mask, int i) {
  return __fns(mask, 0, i + 1);
}

}  // namespace spb
