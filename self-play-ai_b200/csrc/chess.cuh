// chess.cuh — chess rules on bitboards for the device (and the host: every function is __host__ __device__).
//
// Replaces the `State` / `Policy` implementation of the reference's chess adapter, src/game/chess.rs, which delegates
// board mechanics to the crate `chess = "3.2.0"` (Cargo.toml:9; not vendored, so its code is not on this box):
//   get_valid_actions  chess.rs:150-152  -> chess_legal_moves        (bitboard move generation)
//   get_next_state     chess.rs:112-148  -> chess_make_move + the reference's own fifty-move counter (:124-143)
//   get_status         chess.rs:154-166  -> chess_status             (checkmate / stalemate from the move generator,
//                                           legal-move-LIST repetition :51-62, counter >= 100)
//   get_value_and_terminated :168-174    -> chess_terminal_value     (Won = +1.0: the opposite sign of the other games)
//   get_encoding       chess.rs:176-249  -> chess_encode_plane       (19 x 8 x 8)
//   Policy::get_channel / get_action :311-493 -> chess_channel / chess_action (73 move planes)
//
// PARITY UNPINNED by the reference where the crate decides: the ORDER of the legal moves (= the order of a node's
// children, mcts.rs:121-122).  Here moves come out sorted by (from, to, promotion); the set of legal moves is pinned by
// the standard perft counts (tests/test_chess.py) and by an independent array-board restatement (oracle/chess_oracle.cc).
// Squares are rank*8 + file (a1 = 0 ... h8 = 63), as Square::to_index of the crate.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/selfplay_b200.h"

namespace spb {
namespace chess {

enum Piece : int { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, NO_PIECE = 6 };
constexpr int CASTLE_WK = 1, CASTLE_WQ = 2, CASTLE_BK = 4, CASTLE_BQ = 8;
constexpr int NO_SQ = 64;
constexpr int MAX_MOVES = 256;                 // 218 is the known maximum of legal moves in a position
constexpr uint16_t MOVE_NONE = 0xFFFFu;

using Pos = spb_chess_state;                   // plain struct of the C ABI (include/selfplay_b200.h)
using Move = uint16_t;                         // from | to << 6 | promotion << 12 (0 none, else Piece code KNIGHT..QUEEN)

#define SPB_HD __host__ __device__ __forceinline__

SPB_HD Move make_move_code(int from, int to, int promo) { return (Move)(from | (to << 6) | (promo << 12)); }
SPB_HD int move_from(Move m) { return m & 63; }
SPB_HD int move_to(Move m) { return (m >> 6) & 63; }
SPB_HD int move_promo(Move m) { return (m >> 12) & 7; }

SPB_HD int lsb(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)b) - 1;
#else
  return __builtin_ctzll(b);
#endif
}
SPB_HD int popcnt(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __popcll(b);
#else
  return __builtin_popcountll(b);
#endif
}
SPB_HD uint64_t bit(int sq) { return 1ull << sq; }

constexpr uint64_t FILE_A = 0x0101010101010101ull, FILE_H = 0x8080808080808080ull;
constexpr uint64_t FILE_AB = FILE_A | (FILE_A << 1), FILE_GH = FILE_H | (FILE_H >> 1);
constexpr uint64_t RANK_1 = 0xFFull, RANK_2 = 0xFF00ull, RANK_7 = 0xFF000000000000ull, RANK_8 = 0xFF00000000000000ull;

SPB_HD uint64_t knight_attacks(uint64_t b) {
  return ((b << 17) & ~FILE_A) | ((b << 15) & ~FILE_H) | ((b << 10) & ~FILE_AB) | ((b << 6) & ~FILE_GH) |
         ((b >> 17) & ~FILE_H) | ((b >> 15) & ~FILE_A) | ((b >> 10) & ~FILE_GH) | ((b >> 6) & ~FILE_AB);
}
SPB_HD uint64_t king_attacks(uint64_t b) {
  const uint64_t h = ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H);
  const uint64_t r = b | h;
  return h | (r << 8) | (r >> 8);
}
// squares attacked by the pawns `b` of colour c
SPB_HD uint64_t pawn_attacks(uint64_t b, int c) {
  return c == 0 ? (((b << 9) & ~FILE_A) | ((b << 7) & ~FILE_H)) : (((b >> 7) & ~FILE_A) | ((b >> 9) & ~FILE_H));
}
// sliding attacks from sq along one ray (dr, df in {-1,0,1}), stopping at the first occupied square (included)
SPB_HD uint64_t ray(int sq, uint64_t occ, int dr, int df) {
  uint64_t a = 0;
  int r = sq >> 3, f = sq & 7;
  for (;;) {
    r += dr; f += df;
    if ((unsigned)r > 7u || (unsigned)f > 7u) break;
    const uint64_t s = bit(r * 8 + f);
    a |= s;
    if (occ & s) break;
  }
  return a;
}
#ifdef __CUDA_ARCH__
// Device: both rays of one line (file, rank, diagonal, anti-diagonal) at once by the subtraction trick o ^ (o - 2s): the
// borrow of the subtraction runs from the slider to the first blocker; the ray towards lower squares is the same on the
// bit-reversed board (BREV is one instruction).  ~12 instructions per line instead of a loop over squares.
__device__ __forceinline__ uint64_t line_attacks(uint64_t occ, int sq, uint64_t mask) {
  const uint64_t o = occ & mask, s = bit(sq);
  const uint64_t fwd = o - 2 * s;
  const uint64_t rev = __brevll(__brevll(o) - 2 * __brevll(s));
  return (fwd ^ rev) & mask;
}
__device__ __forceinline__ uint64_t rook_attacks(int sq, uint64_t occ) {
  return line_attacks(occ, sq, FILE_A << (sq & 7)) | line_attacks(occ, sq, RANK_1 << (sq & 56));
}
__device__ __forceinline__ uint64_t bishop_attacks(int sq, uint64_t occ) {
  const int d = (sq >> 3) - (sq & 7), a = (sq >> 3) + (sq & 7) - 7;
  const uint64_t MAIN = 0x8040201008040201ull, ANTI = 0x0102040810204080ull;
  const uint64_t dm = d >= 0 ? MAIN << (8 * d) : MAIN >> (8 * -d);
  const uint64_t am = a >= 0 ? ANTI << (8 * a) : ANTI >> (8 * -a);
  return line_attacks(occ, sq, dm) | line_attacks(occ, sq, am);
}
#else
inline uint64_t rook_attacks(int sq, uint64_t occ) { return ray(sq, occ, 1, 0) | ray(sq, occ, -1, 0) | ray(sq, occ, 0, 1) | ray(sq, occ, 0, -1); }
inline uint64_t bishop_attacks(int sq, uint64_t occ) { return ray(sq, occ, 1, 1) | ray(sq, occ, 1, -1) | ray(sq, occ, -1, 1) | ray(sq, occ, -1, -1); }
#endif

SPB_HD uint64_t occupied(const Pos& p) { return p.color[0] | p.color[1]; }

SPB_HD int piece_on(const Pos& p, int sq) {
  const uint64_t s = bit(sq);
#pragma unroll
  for (int t = 0; t < 6; ++t)
    if (p.piece[t] & s) return t;
  return NO_PIECE;
}

// is `sq` attacked by a piece of colour `by`?
SPB_HD bool attacked(const Pos& p, int sq, int by) {
  const uint64_t them = p.color[by], occ = occupied(p), s = bit(sq);
  if (pawn_attacks(p.piece[PAWN] & them, by) & s) return true;
  if (knight_attacks(s) & p.piece[KNIGHT] & them) return true;
  if (king_attacks(s) & p.piece[KING] & them) return true;
  if (rook_attacks(sq, occ) & (p.piece[ROOK] | p.piece[QUEEN]) & them) return true;
  if (bishop_attacks(sq, occ) & (p.piece[BISHOP] | p.piece[QUEEN]) & them) return true;
  return false;
}
SPB_HD bool in_check(const Pos& p, int c) { return attacked(p, lsb(p.piece[KING] & p.color[c]), c ^ 1); }

// Board mechanics of one move (the crate's Board::make_move) + the reference's counters (chess.rs:124-143).  The move
// must be pseudo-legal for `p`.
SPB_HD Pos apply_move(const Pos& p, Move m) {
  Pos q = p;
  const int from = move_from(m), to = move_to(m), promo = move_promo(m);
  const int us = p.side, them = us ^ 1;
  const uint64_t fb = bit(from), tb = bit(to);
  const int pc = piece_on(p, from);
  const int cap = piece_on(p, to);                                 // NO_PIECE for quiet moves and en passant
  // chess.rs:131-134: reversible = not a pawn move, destination empty, castle rights of both colours unchanged
  if (cap != NO_PIECE) { q.piece[cap] &= ~tb; q.color[them] &= ~tb; }
  q.piece[pc] &= ~fb; q.color[us] &= ~fb;
  const int placed = promo ? promo : pc;
  q.piece[placed] |= tb; q.color[us] |= tb;
  if (pc == PAWN && to == p.ep && p.ep != NO_SQ) {                 // en passant: the captured pawn stands beside the target
    const uint64_t cb = bit(us == 0 ? to - 8 : to + 8);
    q.piece[PAWN] &= ~cb; q.color[them] &= ~cb;
  }
  if (pc == KING && (to - from == 2 || from - to == 2)) {          // castling: the rook jumps over the king
    const int rf = to > from ? from + 3 : from - 4, rt = to > from ? from + 1 : from - 1;
    q.piece[ROOK] &= ~bit(rf); q.color[us] &= ~bit(rf);
    q.piece[ROOK] |= bit(rt); q.color[us] |= bit(rt);
  }
  uint8_t cr = p.castle;
  if (pc == KING) cr &= us == 0 ? ~(CASTLE_WK | CASTLE_WQ) : ~(CASTLE_BK | CASTLE_BQ);
  if (from == 0 || to == 0) cr &= ~CASTLE_WQ;
  if (from == 7 || to == 7) cr &= ~CASTLE_WK;
  if (from == 56 || to == 56) cr &= ~CASTLE_BQ;
  if (from == 63 || to == 63) cr &= ~CASTLE_BK;
  q.castle = cr;
  q.ep = NO_SQ;
  if (pc == PAWN && (to - from == 16 || from - to == 16)) {
    const int mid = (from + to) >> 1;
    if (pawn_attacks(bit(mid), us) & p.piece[PAWN] & p.color[them]) q.ep = (uint8_t)mid;   // only when a capture is possible
  }
  const bool reversible = pc != PAWN && cap == NO_PIECE && cr == p.castle;
  q.fifty = reversible ? (uint16_t)(p.fifty + 1) : (uint16_t)0;
  q.plies = (uint16_t)(p.plies + 1);
  q.side = (uint8_t)them;
  return q;
}

// get_valid_actions (chess.rs:150-152): the legal moves of the side to move, sorted by (from, to, promotion).  Returns
// their number; `out` needs MAX_MOVES entries.
SPB_HD int legal_moves(const Pos& p, Move* out) {
  int n = 0;
  const int us = p.side, them = us ^ 1;
  const uint64_t own = p.color[us], enemy = p.color[them], occ = own | enemy;
  const int ksq = lsb(p.piece[KING] & own);
  uint64_t pieces = own;
  while (pieces) {
    const int from = lsb(pieces);
    pieces &= pieces - 1;
    const uint64_t fb = bit(from);
    const int pc = piece_on(p, from);
    uint64_t targets = 0;
    switch (pc) {
      case PAWN: {
        if (us == 0) {
          const uint64_t one = (fb << 8) & ~occ;
          targets = one | (((one & (RANK_2 << 8)) << 8) & ~occ);
        } else {
          const uint64_t one = (fb >> 8) & ~occ;
          targets = one | (((one & (RANK_7 >> 8)) >> 8) & ~occ);
        }
        targets |= pawn_attacks(fb, us) & (enemy | (p.ep != NO_SQ ? bit(p.ep) : 0ull));
        break;
      }
      case KNIGHT: targets = knight_attacks(fb) & ~own; break;
      case BISHOP: targets = bishop_attacks(from, occ) & ~own; break;
      case ROOK: targets = rook_attacks(from, occ) & ~own; break;
      case QUEEN: targets = (rook_attacks(from, occ) | bishop_attacks(from, occ)) & ~own; break;
      default: {
        targets = king_attacks(fb) & ~own;
        // castling: rights, empty squares between, king not in check and not passing over an attacked square
        const int ks = us == 0 ? CASTLE_WK : CASTLE_BK, qs = us == 0 ? CASTLE_WQ : CASTLE_BQ;
        if ((p.castle & ks) && !(occ & (bit(from + 1) | bit(from + 2))) && !attacked(p, from, them) && !attacked(p, from + 1, them))
          targets |= bit(from + 2);
        if ((p.castle & qs) && !(occ & (bit(from - 1) | bit(from - 2) | bit(from - 3))) && !attacked(p, from, them) && !attacked(p, from - 1, them))
          targets |= bit(from - 2);
        break;
      }
    }
    while (targets) {
      const int to = lsb(targets);
      targets &= targets - 1;
      const bool promotes = pc == PAWN && (bit(to) & (RANK_1 | RANK_8));
      const Move probe = make_move_code(from, to, promotes ? QUEEN : 0);
      const Pos q = apply_move(p, probe);
      const int k2 = pc == KING ? to : ksq;
      if (attacked(q, k2, them)) continue;                         // leaves the own king in check
      if (promotes) {
        for (int pr = KNIGHT; pr <= QUEEN; ++pr) out[n++] = make_move_code(from, to, pr);
      } else {
        out[n++] = probe;
      }
    }
  }
  return n;
}

// Hash of a legal-move list (number of moves and every (index, move) pair): the reference's repetition rule compares the
// legal move LISTS of positions (chess.rs:51-62), so the per-ply history holds these hashes.  A wrapping sum of mixed
// (index, move) words, so that a warp can hash its list lane-parallel (chess_tree.cuh) and still get this value.
SPB_HD uint64_t mix64(uint64_t x) {                                 // splitmix64 finaliser (games.cuh has the same function)
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
SPB_HD uint64_t move_hash_term(int index, Move m) { return mix64(((uint64_t)index << 16) | (uint64_t)m); }
SPB_HD uint64_t move_list_hash_seed(int n) { return (uint64_t)n * 0x9E3779B97F4A7C15ull; }
SPB_HD uint64_t move_list_hash(const Move* mv, int n) {
  uint64_t h = move_list_hash_seed(n);
  for (int i = 0; i < n; ++i) h += move_hash_term(i, mv[i]);
  return h;
}

// get_num_repetitions, chess.rs:51-62: 1 + number of earlier positions of the game whose legal move list equals the
// current one.  history[i] = move_list_hash of the position before ply i (transposition_table, chess.rs:121-122).
SPB_HD uint32_t num_repetitions(uint64_t current_hash, const uint64_t* history, uint32_t hist_len) {
  uint32_t c = 0;
  for (uint32_t i = 0; i < hist_len; ++i) c += history[i] == current_hash;
  return c + 1;
}

// get_status, chess.rs:154-166.
SPB_HD uint32_t status(const Pos& p, int n_legal, uint64_t current_hash, const uint64_t* history, uint32_t hist_len) {
  if (n_legal == 0) return in_check(p, p.side) ? SPB_STATUS_WON : SPB_STATUS_TIED;   // checkmate / stalemate
  if (num_repetitions(current_hash, history, hist_len) >= 3 || p.fifty >= 100) return SPB_STATUS_TIED;
  return SPB_STATUS_ONGOING;
}
// get_value_and_terminated, chess.rs:168-174: Won is +1.0 here (connect_four.rs:236 and tictactoe.rs:193 use -1.0).
SPB_HD float terminal_value(uint32_t st) { return st == SPB_STATUS_WON ? 1.0f : 0.0f; }

// get_encoding, chess.rs:176-249: value of encoding[plane][row][col].  Rows are ranks seen from the side to move (black:
// rank 7 - row), files are NOT mirrored.  Planes 0..5 own P N B R Q K, 6..11 the opponent's, 12/13 own king/queen-side
// castling right, 14/15 the opponent's, 16 repetitions, 17 fifty-move counter / 100, 18 (moves played / 2) / 50.
SPB_HD float encode_plane(const Pos& p, uint32_t repetitions, int plane, int row, int col) {
  const int us = p.side;
  if (plane < 12) {
    const int rank = us == 0 ? row : 7 - row;
    const uint64_t s = bit(rank * 8 + col);
    const int t = plane % 6, c = plane < 6 ? us : us ^ 1;
    return (p.piece[t] & p.color[c] & s) ? 1.0f : 0.0f;
  }
  const int mk = us == 0 ? CASTLE_WK : CASTLE_BK, mq = us == 0 ? CASTLE_WQ : CASTLE_BQ;
  const int tk = us == 0 ? CASTLE_BK : CASTLE_WK, tq = us == 0 ? CASTLE_BQ : CASTLE_WQ;
  switch (plane) {
    case 12: return (p.castle & mk) ? 1.0f : 0.0f;
    case 13: return (p.castle & mq) ? 1.0f : 0.0f;
    case 14: return (p.castle & tk) ? 1.0f : 0.0f;
    case 15: return (p.castle & tq) ? 1.0f : 0.0f;
    case 16: return (float)repetitions;
    case 17: return (float)p.fifty / 100.0f;
    default: return (float)(p.plies / 2) / 50.0f;
  }
}

// ---- 73 move planes (chess.rs:8-18) ----------------------------------------------------------------------------
constexpr int ROOK_PROMO0 = 0, BISHOP_PROMO0 = 3, KNIGHT_PROMO0 = 6, HORIZONTAL0 = 9, VERTICAL0 = 23, DIAGONAL0 = 37, KNIGHT0 = 65;

// Policy::get_channel, chess.rs:311-390: the plane of a move for the player `side` (rank differences are seen from the
// mover's side; file differences are not mirrored).
SPB_HD int channel(int side, Move m) {
  const int from = move_from(m), to = move_to(m), promo = move_promo(m);
  int rank_diff = (to >> 3) - (from >> 3);
  const int file_diff = (to & 7) - (from & 7);
  const int ard = rank_diff < 0 ? -rank_diff : rank_diff, afd = file_diff < 0 ? -file_diff : file_diff;
  if (side == 1) rank_diff = -rank_diff;
  const int sub = file_diff + 1;                                   // left 0, straight 1, right 2
  if (promo == ROOK) return ROOK_PROMO0 + sub;
  if (promo == BISHOP) return BISHOP_PROMO0 + sub;
  if (promo == KNIGHT) return KNIGHT_PROMO0 + sub;
  if (rank_diff == 0) return file_diff < 0 ? HORIZONTAL0 + (-file_diff) - 1 : HORIZONTAL0 + 7 + file_diff - 1;
  if (file_diff == 0) return rank_diff < 0 ? VERTICAL0 + (-rank_diff) - 1 : VERTICAL0 + 7 + rank_diff - 1;
  if (ard == afd) {
    if (file_diff < 0) return rank_diff > 0 ? DIAGONAL0 + rank_diff - 1 : DIAGONAL0 + 7 + (-rank_diff) - 1;
    return rank_diff > 0 ? DIAGONAL0 + 14 + rank_diff - 1 : DIAGONAL0 + 21 + (-rank_diff) - 1;
  }
  if (file_diff < 0) {
    if (rank_diff > 0) return ard > afd ? KNIGHT0 : KNIGHT0 + 1;
    return ard > afd ? KNIGHT0 + 2 : KNIGHT0 + 3;
  }
  if (rank_diff > 0) return ard > afd ? KNIGHT0 + 4 : KNIGHT0 + 5;
  return ard > afd ? KNIGHT0 + 6 : KNIGHT0 + 7;
}
// Index of a move in the flat 73*8*8 policy (get_prob / set_prob, chess.rs:495-514): [channel][row][file of the source],
// row = source rank seen from the mover.
SPB_HD int policy_index(int side, Move m) {
  const int from = move_from(m);
  const int row = side == 1 ? 7 - (from >> 3) : (from >> 3);
  return channel(side, m) * 64 + row * 8 + (from & 7);
}
// Policy::get_action, chess.rs:392-493, LITERALLY — including its slip at :442, which compares against
// KNIGHT_MOVE_START_IDX where KNIGHT_PROMOTION_START_IDX is meant, so that the knight-promotion planes 6..8 decode to
// file differences +2..+4.  The function is off the search path (only sample / get_best_action use it).
// Returns MOVE_NONE when a square falls off the board.
SPB_HD Move action(int side, int ch, int row, int col) {
  int promo = 0;
  if (ch < BISHOP_PROMO0) promo = ROOK; else if (ch < KNIGHT_PROMO0) promo = BISHOP; else if (ch < HORIZONTAL0) promo = KNIGHT;
  int rank_diff, file_diff;
  if (ch < HORIZONTAL0) rank_diff = 1;
  else if (ch < VERTICAL0) rank_diff = 0;
  else if (ch < DIAGONAL0) { const int o = ch - VERTICAL0; rank_diff = o < 7 ? -(o + 1) : o + 1 - 7; }
  else if (ch < KNIGHT0) {
    const int o = ch - DIAGONAL0;
    rank_diff = o < 7 ? o + 1 : (o < 14 ? -(o + 1 - 7) : (o < 21 ? o + 1 - 14 : -(o + 1 - 21)));
  } else {
    const int o = ch - KNIGHT0;
    rank_diff = (o == 0 || o == 4) ? 2 : ((o == 1 || o == 5) ? 1 : ((o == 2 || o == 6) ? -2 : -1));
  }
  if (ch < HORIZONTAL0) {
    if (ch < BISHOP_PROMO0) file_diff = ch - 1;
    else if (ch < KNIGHT0) file_diff = ch - BISHOP_PROMO0 - 1;     // chess.rs:442 (sic): also taken by the knight promotions
    else file_diff = ch - KNIGHT_PROMO0 - 1;
  } else if (ch < VERTICAL0) { const int o = ch - HORIZONTAL0; file_diff = o < 7 ? -(o + 1) : o + 1 - 7; }
  else if (ch < DIAGONAL0) file_diff = 0;
  else if (ch < KNIGHT0) {
    const int o = ch - DIAGONAL0;
    file_diff = o < 7 ? -(o + 1) : (o < 14 ? -(o + 1 - 7) : (o < 21 ? o + 1 - 14 : o + 1 - 21));
  } else {
    const int o = ch - KNIGHT0;
    file_diff = (o == 0 || o == 2) ? -1 : ((o == 1 || o == 3) ? -2 : ((o == 4 || o == 6) ? 1 : 2));
  }
  if (side == 1) { rank_diff = -rank_diff; row = 7 - row; }
  const int r2 = row + rank_diff, c2 = col + file_diff;
  if ((unsigned)r2 > 7u || (unsigned)c2 > 7u) return MOVE_NONE;
  return make_move_code(row * 8 + col, r2 * 8 + c2, promo);
}

SPB_HD Pos start_position() {
  Pos p{};
  p.piece[PAWN] = RANK_2 | RANK_7;
  p.piece[KNIGHT] = 0x4200000000000042ull;
  p.piece[BISHOP] = 0x2400000000000024ull;
  p.piece[ROOK] = 0x8100000000000081ull;
  p.piece[QUEEN] = 0x0800000000000008ull;
  p.piece[KING] = 0x1000000000000010ull;
  p.color[0] = 0xFFFFull;
  p.color[1] = 0xFFFF000000000000ull;
  p.side = 0;
  p.castle = CASTLE_WK | CASTLE_WQ | CASTLE_BK | CASTLE_BQ;
  p.ep = NO_SQ;
  p.fifty = 0;
  p.plies = 0;
  return p;
}

#undef SPB_HD
}  // namespace chess
}  // namespace spb
