/*
 * chess_oracle.cc — CPU restatement of the reference's chess adapter (src/game/chess.rs).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/oracle.h).  PARITY UNPINNED BY THE REFERENCE: chess.rs delegates the
 * board mechanics and the ORDER of the legal moves to the crate `chess = "3.2.0"` (Cargo.toml:9), which is not
 * vendored; the reference ships no test, fixture or golden vector.  What pins this file instead:
 *   - the published perft counts of the standard test positions (start position, "Kiwipete", positions 3-6 of the
 *     chessprogramming wiki) pin the legal-move SETS: castling, en passant, promotions, pins, checks;
 *   - everything chess.rs itself computes is restated line by line with its citation: the repetition rule on legal-move
 *     LISTS (:51-62, :121-122), the reversible-move counter (:124-143), get_status (:154-166), the +1.0 terminal value
 *     (:168-174), the 19x8x8 encoding (:176-249), the 73 move planes get_channel / get_action (:311-493, including the
 *     slip at :442).
 * The implementation is deliberately unlike the device's (csrc/chess.cuh, bitboards): a 64-entry array board, move
 * generation by walking offsets, repetition by comparing whole move lists.  Legal moves are listed sorted by
 * (from, to, promotion) — the order this repo defines in place of the crate's.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/selfplay_b200.h"

namespace {

enum { EMPTY = 0, WP = 1, WN, WB, WR, WQ, WK, BP, BN, BB, BR, BQ, BK };
typedef uint16_t Move;   // from | to << 6 | promo << 12, promo: 0 none, 1 knight, 2 bishop, 3 rook, 4 queen

inline int color_of(int pc) { return pc == EMPTY ? -1 : (pc >= BP ? 1 : 0); }
inline int type_of(int pc) { return pc == EMPTY ? -1 : (pc - 1) % 6; }   // 0 pawn .. 5 king
inline Move mk(int f, int t, int p) { return (Move)(f | (t << 6) | (p << 12)); }

struct Board {
  int sq[64];
  int side = 0;
  int castle = 15;      // 1 WK, 2 WQ, 4 BK, 8 BQ
  int ep = 64;
  int fifty = 0;        // chess.rs:28 fifty_move_rule_halfmove_counter
  int plies = 0;        // game.actions() MakeMove count
};

struct Game {
  Board b;
  std::vector<std::vector<Move>> table;   // chess.rs:27 transposition_table: legal move list of the position before each ply
};

bool on_board(int r, int f) { return r >= 0 && r < 8 && f >= 0 && f < 8; }

bool attacked(const Board& b, int s, int by) {
  const int r = s / 8, f = s % 8;
  // pawns
  const int pr = by == 0 ? r - 1 : r + 1;
  for (int df = -1; df <= 1; df += 2)
    if (on_board(pr, f + df) && b.sq[pr * 8 + f + df] == (by == 0 ? WP : BP)) return true;
  static const int kn[8][2] = {{1, 2}, {2, 1}, {-1, 2}, {-2, 1}, {1, -2}, {2, -1}, {-1, -2}, {-2, -1}};
  for (auto& d : kn)
    if (on_board(r + d[0], f + d[1]) && b.sq[(r + d[0]) * 8 + f + d[1]] == (by == 0 ? WN : BN)) return true;
  for (int dr = -1; dr <= 1; ++dr)
    for (int df = -1; df <= 1; ++df) {
      if (!dr && !df) continue;
      if (on_board(r + dr, f + df) && b.sq[(r + dr) * 8 + f + df] == (by == 0 ? WK : BK)) return true;
      int rr = r + dr, ff = f + df;
      while (on_board(rr, ff)) {
        const int pc = b.sq[rr * 8 + ff];
        if (pc != EMPTY) {
          if (color_of(pc) == by) {
            const int t = type_of(pc);
            const bool diag = dr != 0 && df != 0;
            if (t == 4 || (diag && t == 2) || (!diag && t == 3)) return true;
          }
          break;
        }
        rr += dr; ff += df;
      }
    }
  return false;
}

int king_square(const Board& b, int c) {
  for (int s = 0; s < 64; ++s)
    if (b.sq[s] == (c == 0 ? WK : BK)) return s;
  return -1;
}

Board apply(const Board& b, Move m) {
  Board q = b;
  const int from = m & 63, to = (m >> 6) & 63, promo = (m >> 12) & 7;
  const int pc = b.sq[from], cap = b.sq[to];
  const int us = b.side;
  q.sq[from] = EMPTY;
  q.sq[to] = promo ? (us == 0 ? WP : BP) + promo : pc;
  if (type_of(pc) == 0 && to == b.ep && b.ep != 64) q.sq[us == 0 ? to - 8 : to + 8] = EMPTY;
  if (type_of(pc) == 5 && std::abs(to - from) == 2) {
    const int rf = to > from ? from + 3 : from - 4, rt = to > from ? from + 1 : from - 1;
    q.sq[rt] = q.sq[rf];
    q.sq[rf] = EMPTY;
  }
  if (type_of(pc) == 5) q.castle &= us == 0 ? ~3 : ~12;
  if (from == 0 || to == 0) q.castle &= ~2;
  if (from == 7 || to == 7) q.castle &= ~1;
  if (from == 56 || to == 56) q.castle &= ~8;
  if (from == 63 || to == 63) q.castle &= ~4;
  q.ep = 64;
  if (type_of(pc) == 0 && std::abs(to - from) == 16) {
    const int mid = (from + to) / 2, r = mid / 8, f = mid % 8, pr = us == 0 ? r + 1 : r - 1;
    for (int df = -1; df <= 1; df += 2)
      if (on_board(pr, f + df) && b.sq[pr * 8 + f + df] == (us == 0 ? BP : WP)) q.ep = mid;
  }
  // chess.rs:131-134: reversible <=> not a pawn move, destination square empty, castle rights of both sides unchanged
  const bool reversible = type_of(pc) != 0 && cap == EMPTY && q.castle == b.castle;
  q.fifty = reversible ? b.fifty + 1 : 0;                            // chess.rs:139-144
  q.plies = b.plies + 1;
  q.side = us ^ 1;
  return q;
}

void add_if_legal(const Board& b, int from, int to, std::vector<Move>& out) {
  const int pc = b.sq[from], us = b.side;
  const bool promotes = type_of(pc) == 0 && (to / 8 == 0 || to / 8 == 7);
  const Board q = apply(b, mk(from, to, promotes ? 4 : 0));
  if (attacked(q, king_square(q, us), us ^ 1)) return;
  if (promotes) {
    for (int p = 1; p <= 4; ++p) out.push_back(mk(from, to, p));
  } else {
    out.push_back(mk(from, to, 0));
  }
}

// MoveGen::new_legal(&board).collect() (chess.rs:151), in this repo's canonical order.
std::vector<Move> legal_moves(const Board& b) {
  std::vector<Move> out;
  const int us = b.side, them = us ^ 1;
  for (int from = 0; from < 64; ++from) {
    const int pc = b.sq[from];
    if (color_of(pc) != us) continue;
    const int r = from / 8, f = from % 8, t = type_of(pc);
    std::vector<int> targets;
    if (t == 0) {
      const int dir = us == 0 ? 1 : -1, start = us == 0 ? 1 : 6;
      if (on_board(r + dir, f) && b.sq[(r + dir) * 8 + f] == EMPTY) {
        targets.push_back((r + dir) * 8 + f);
        if (r == start && b.sq[(r + 2 * dir) * 8 + f] == EMPTY) targets.push_back((r + 2 * dir) * 8 + f);
      }
      for (int df = -1; df <= 1; df += 2) {
        if (!on_board(r + dir, f + df)) continue;
        const int to = (r + dir) * 8 + f + df;
        if (color_of(b.sq[to]) == them || (to == b.ep && b.ep != 64)) targets.push_back(to);
      }
    } else if (t == 1) {
      static const int kn[8][2] = {{1, 2}, {2, 1}, {-1, 2}, {-2, 1}, {1, -2}, {2, -1}, {-1, -2}, {-2, -1}};
      for (auto& d : kn)
        if (on_board(r + d[0], f + d[1]) && color_of(b.sq[(r + d[0]) * 8 + f + d[1]]) != us) targets.push_back((r + d[0]) * 8 + f + d[1]);
    } else if (t == 5) {
      for (int dr = -1; dr <= 1; ++dr)
        for (int df = -1; df <= 1; ++df)
          if ((dr || df) && on_board(r + dr, f + df) && color_of(b.sq[(r + dr) * 8 + f + df]) != us) targets.push_back((r + dr) * 8 + f + df);
      const int ks = us == 0 ? 1 : 4, qs = us == 0 ? 2 : 8;
      if ((b.castle & ks) && b.sq[from + 1] == EMPTY && b.sq[from + 2] == EMPTY && !attacked(b, from, them) && !attacked(b, from + 1, them))
        targets.push_back(from + 2);
      if ((b.castle & qs) && b.sq[from - 1] == EMPTY && b.sq[from - 2] == EMPTY && b.sq[from - 3] == EMPTY && !attacked(b, from, them) &&
          !attacked(b, from - 1, them))
        targets.push_back(from - 2);
    } else {
      for (int dr = -1; dr <= 1; ++dr)
        for (int df = -1; df <= 1; ++df) {
          if (!dr && !df) continue;
          const bool diag = dr != 0 && df != 0;
          if ((t == 2 && !diag) || (t == 3 && diag)) continue;
          int rr = r + dr, ff = f + df;
          while (on_board(rr, ff)) {
            const int c = color_of(b.sq[rr * 8 + ff]);
            if (c != us) targets.push_back(rr * 8 + ff);
            if (c != -1) break;
            rr += dr; ff += df;
          }
        }
    }
    std::sort(targets.begin(), targets.end());
    for (int to : targets) add_if_legal(b, from, to, out);
  }
  return out;
}

// chess.rs:51-62
int num_repetitions(const Game& g) {
  const std::vector<Move> cur = legal_moves(g.b);
  int counter = 0;
  for (const auto& pos : g.table)
    if (pos == cur) ++counter;
  return counter + 1;
}

// chess.rs:154-166
int status(const Game& g) {
  const std::vector<Move> mv = legal_moves(g.b);
  if (mv.empty()) return attacked(g.b, king_square(g.b, g.b.side), g.b.side ^ 1) ? SPB_STATUS_WON : SPB_STATUS_TIED;
  if (num_repetitions(g) >= 3 || g.b.fifty >= 100) return SPB_STATUS_TIED;
  return SPB_STATUS_ONGOING;
}

bool parse_fen(const char* fen, Board* b) {
  std::memset(b->sq, 0, sizeof b->sq);
  int r = 7, f = 0;
  const char* p = fen;
  for (; *p && *p != ' '; ++p) {
    if (*p == '/') { --r; f = 0; continue; }
    if (*p >= '1' && *p <= '8') { f += *p - '0'; continue; }
    static const char* names = "PNBRQKpnbrqk";
    const char* q = std::strchr(names, *p);
    if (!q || r < 0 || f > 7) return false;
    b->sq[r * 8 + f++] = 1 + (int)(q - names);
  }
  if (*p != ' ') return false;
  ++p;
  b->side = *p == 'b';
  p += 2;
  b->castle = 0;
  for (; *p && *p != ' '; ++p) {
    if (*p == 'K') b->castle |= 1;
    if (*p == 'Q') b->castle |= 2;
    if (*p == 'k') b->castle |= 4;
    if (*p == 'q') b->castle |= 8;
  }
  if (*p == ' ') ++p;
  b->ep = 64;
  if (*p && *p != '-') {
    const int mid = (p[1] - '1') * 8 + (p[0] - 'a');
    // kept only when a capture is possible, as after a move (see apply)
    const int us = b->side, rr = mid / 8 + (us == 0 ? -1 : 1);
    for (int df = -1; df <= 1; df += 2)
      if (on_board(rr, mid % 8 + df) && b->sq[rr * 8 + mid % 8 + df] == (us == 0 ? WP : BP)) b->ep = mid;
  }
  b->fifty = 0;
  b->plies = 0;
  // optional halfmove clock (start value of the reference's reversible-move counter) and fullmove number
  while (*p && *p != ' ') ++p;
  if (*p == ' ') {
    char* end = nullptr;
    const long half = std::strtol(p, &end, 10);
    if (end != p) {
      b->fifty = (int)half;
      const long full = std::strtol(end, &end, 10);
      if (full >= 1) b->plies = (int)(full - 1) * 2 + b->side;
    }
  }
  return true;
}

uint64_t perft(const Board& b, int depth) {
  const std::vector<Move> mv = legal_moves(b);
  if (depth <= 1) return depth == 1 ? mv.size() : 1;
  uint64_t n = 0;
  for (Move m : mv) n += perft(apply(b, m), depth - 1);
  return n;
}

// the list hash of csrc/chess.cuh (move_list_hash), restated: the device keeps one hash per ply where this file keeps lists
uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t list_hash(const std::vector<Move>& mv) {
  uint64_t h = (uint64_t)mv.size() * 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < mv.size(); ++i) h += mix64(((uint64_t)i << 16) | (uint64_t)mv[i]);
  return h;
}

}  // namespace

extern "C" {

struct orc_chess {
  Game g;
};

orc_chess* orc_chess_new(const char* fen) {
  orc_chess* c = new orc_chess();
  if (!parse_fen(fen ? fen : "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", &c->g.b)) { delete c; return nullptr; }
  return c;
}
void orc_chess_free(orc_chess* c) { delete c; }
orc_chess* orc_chess_clone(const orc_chess* c) { return new orc_chess(*c); }

int32_t orc_chess_legal_moves(const orc_chess* c, uint16_t* out) {
  const std::vector<Move> mv = legal_moves(c->g.b);
  if (out) std::copy(mv.begin(), mv.end(), out);
  return (int32_t)mv.size();
}

// get_next_state, chess.rs:112-148
int32_t orc_chess_make_move(orc_chess* c, uint16_t m) {
  if (status(c->g) != SPB_STATUS_ONGOING) return SPB_ERR_ILLEGAL;   // :113-115 "Game is already over"
  const std::vector<Move> mv = legal_moves(c->g.b);
  if (std::find(mv.begin(), mv.end(), (Move)m) == mv.end()) return SPB_ERR_ILLEGAL;   // :146 "Failed to make move"
  c->g.table.push_back(mv);                                         // :121-122
  c->g.b = apply(c->g.b, (Move)m);
  return SPB_OK;
}

int32_t orc_chess_status(const orc_chess* c) { return status(c->g); }
int32_t orc_chess_repetitions(const orc_chess* c) { return num_repetitions(c->g); }
// get_value_and_terminated, chess.rs:168-174
float orc_chess_value(const orc_chess* c) { return status(c->g) == SPB_STATUS_WON ? 1.0f : 0.0f; }
int32_t orc_chess_side(const orc_chess* c) { return c->g.b.side; }
uint64_t orc_chess_perft(const orc_chess* c, int32_t depth) { return perft(c->g.b, depth); }

// get_encoding, chess.rs:176-249: out[19][8][8]
void orc_chess_encode(const orc_chess* c, float* out) {
  const Board& b = c->g.b;
  std::memset(out, 0, sizeof(float) * 19 * 64);
  const int me = b.side;
  for (int row = 0; row < 8; ++row) {
    const int rank = me == 0 ? row : 7 - row;                       // :184-187
    for (int col = 0; col < 8; ++col) {
      const int pc = b.sq[rank * 8 + col];
      if (pc == EMPTY) continue;
      const int offset = color_of(pc) == me ? 0 : 6;                // :193-201
      out[(offset + type_of(pc)) * 64 + row * 8 + col] = 1.0f;     // :203-211
    }
  }
  auto fill = [&](int plane, float v) { for (int i = 0; i < 64; ++i) out[plane * 64 + i] = v; };
  const int mk_ = me == 0 ? 1 : 4, mq = me == 0 ? 2 : 8, tk = me == 0 ? 4 : 1, tq = me == 0 ? 8 : 2;
  if (b.castle & mk_) fill(12, 1.0f);                               // :216-222
  if (b.castle & mq) fill(13, 1.0f);
  if (b.castle & tk) fill(14, 1.0f);                                // :223-229
  if (b.castle & tq) fill(15, 1.0f);
  fill(16, (float)num_repetitions(c->g));                           // :232
  fill(17, (float)b.fifty / 100.0f);                                // :236
  fill(18, (float)(b.plies / 2) / 50.0f);                           // :240-244
}

// Policy::get_channel, chess.rs:311-390
int32_t orc_chess_channel(int32_t player, uint16_t m) {
  const int from = m & 63, to = (m >> 6) & 63, promo = (m >> 12) & 7;
  long rank_diff = (long)(to / 8) - (long)(from / 8);
  const long file_diff = (long)(to % 8) - (long)(from % 8);
  const long abs_rank_diff = std::labs(rank_diff), abs_file_diff = std::labs(file_diff);
  if (player == 1) rank_diff *= -1;                                 // :322-324
  const long sub_idx = file_diff + 1;                               // :327
  if (promo == 3) return (int32_t)(0 + sub_idx);                    // rook   :329
  if (promo == 2) return (int32_t)(3 + sub_idx);                    // bishop :330
  if (promo == 1) return (int32_t)(6 + sub_idx);                    // knight :331
  if (rank_diff == 0) return (int32_t)(file_diff < 0 ? 9 + (-file_diff) - 1 : 9 + 7 + file_diff - 1);            // :336-341
  if (file_diff == 0) return (int32_t)(rank_diff < 0 ? 23 + (-rank_diff) - 1 : 23 + 7 + rank_diff - 1);           // :344-349
  if (abs_rank_diff == abs_file_diff) {                             // :352-364
    if (file_diff < 0) return (int32_t)(rank_diff > 0 ? 37 + rank_diff - 1 : 37 + 7 + (-rank_diff) - 1);
    return (int32_t)(rank_diff > 0 ? 37 + 14 + rank_diff - 1 : 37 + 21 + (-rank_diff) - 1);
  }
  if (file_diff < 0) {                                              // :367-379
    if (rank_diff > 0) return abs_rank_diff > abs_file_diff ? 65 : 66;
    return abs_rank_diff > abs_file_diff ? 67 : 68;
  }
  if (rank_diff > 0) return abs_rank_diff > abs_file_diff ? 69 : 70;   // :380-384
  return abs_rank_diff > abs_file_diff ? 71 : 72;                   // :385-389
}

// Policy::get_action, chess.rs:392-493 (0xFFFF when a square is off the board — the Rust would build an invalid Square)
uint16_t orc_chess_action(int32_t player, int32_t channel, int32_t row, int32_t col) {
  const int ROOK0 = 0, BISHOP0 = 3, KNIGHTP0 = 6, HOR0 = 9, VER0 = 23, DIA0 = 37, KN0 = 65, STEPS = 7;
  (void)ROOK0;
  int promo = 0;
  if (channel < BISHOP0) promo = 3; else if (channel < KNIGHTP0) promo = 2; else if (channel < HOR0) promo = 1;   // :393-402
  long rank_diff;
  if (channel < HOR0) rank_diff = 1;                                // :405-406
  else if (channel < VER0) rank_diff = 0;
  else if (channel < DIA0) {
    const int offset = channel - VER0;
    rank_diff = offset < STEPS ? -(long)(channel + 1 - VER0) : (long)(channel + 1 - VER0 - STEPS);
  } else if (channel < KN0) {
    const int offset = channel - DIA0;
    if (offset < STEPS) rank_diff = channel + 1 - DIA0;
    else if (offset < 2 * STEPS) rank_diff = -(long)(channel + 1 - DIA0 - STEPS);
    else if (offset < 3 * STEPS) rank_diff = channel + 1 - DIA0 - 2 * STEPS;
    else rank_diff = -(long)(channel + 1 - DIA0 - 3 * STEPS);
  } else {
    switch (channel - KN0) { case 0: case 4: rank_diff = 2; break; case 1: case 5: rank_diff = 1; break; case 2: case 6: rank_diff = -2; break; default: rank_diff = -1; }
  }
  long file_diff;
  if (channel < HOR0) {                                             // :439-446
    if (channel < BISHOP0) file_diff = channel - 1;
    else if (channel < KN0) file_diff = channel - BISHOP0 - 1;      // :442 compares with KNIGHT_MOVE_START_IDX (sic)
    else file_diff = channel - KNIGHTP0 - 1;
  } else if (channel < VER0) {
    const int offset = channel - HOR0;
    file_diff = offset < STEPS ? -(long)(channel + 1 - HOR0) : (long)(channel + 1 - HOR0 - STEPS);
  } else if (channel < DIA0) file_diff = 0;
  else if (channel < KN0) {
    const int offset = channel - DIA0;
    if (offset < STEPS) file_diff = -(long)(channel + 1 - DIA0);
    else if (offset < 2 * STEPS) file_diff = -(long)(channel + 1 - DIA0 - STEPS);
    else if (offset < 3 * STEPS) file_diff = channel + 1 - DIA0 - 2 * STEPS;
    else file_diff = channel + 1 - DIA0 - 3 * STEPS;
  } else {
    switch (channel - KN0) { case 0: case 2: file_diff = -1; break; case 1: case 3: file_diff = -2; break; case 4: case 6: file_diff = 1; break; default: file_diff = 2; }
  }
  if (player == 1) { rank_diff *= -1; row = 7 - row; }              // :474-477
  const long r2 = row + rank_diff, c2 = col + file_diff;
  if (r2 < 0 || r2 > 7 || c2 < 0 || c2 > 7) return 0xFFFFu;
  return mk(row * 8 + col, (int)(r2 * 8 + c2), promo);
}

// Position + per-ply history in the C ABI's form (spb_chess_state, hashes of the legal-move lists).
void orc_chess_export(const orc_chess* c, spb_chess_state* out, uint64_t* history) {
  const Board& b = c->g.b;
  std::memset(out, 0, sizeof *out);
  for (int s = 0; s < 64; ++s) {
    const int pc = b.sq[s];
    if (pc == EMPTY) continue;
    out->piece[type_of(pc)] |= 1ull << s;
    out->color[color_of(pc)] |= 1ull << s;
  }
  out->side = (uint8_t)b.side;
  out->castle = (uint8_t)b.castle;
  out->ep = (uint8_t)b.ep;
  out->fifty = (uint16_t)b.fifty;
  out->plies = (uint16_t)b.plies;
  out->hist_len = (uint32_t)c->g.table.size();
  if (history)
    for (size_t i = 0; i < c->g.table.size() && i < SPB_CHESS_MAX_HISTORY; ++i) history[i] = list_hash(c->g.table[i]);
}

}  // extern "C"

// =====================================================================================================================
// MCTS over chess states: src/mcts.rs restated for chess.rs's State (the small games have theirs in oracle.cc)
// =====================================================================================================================
namespace {

// ndarray 0.15.6 `sum()` on a contiguous f32 array (numeric_util::unrolled_fold), as in oracle.cc; call site chess.rs:268
float ndarray_sum(const float* xs, size_t n) {
  float acc = 0.0f;
  float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  while (n >= 8) {
    for (int i = 0; i < 8; ++i) p[i] = p[i] + xs[i];
    xs += 8;
    n -= 8;
  }
  acc = acc + (p[0] + p[4]);
  acc = acc + (p[1] + p[5]);
  acc = acc + (p[2] + p[6]);
  acc = acc + (p[3] + p[7]);
  for (size_t i = 0; i < n && i < 7; ++i) acc = acc + xs[i];
  return acc;
}

constexpr int POLICY = SPB_CHESS_POLICY_SIZE;

// Policy::get_prob / set_prob index, chess.rs:495-514
int policy_index(int player, Move m) {
  const int from = m & 63;
  int row = from / 8;
  if (player == 1) row = 7 - row;
  return orc_chess_channel(player, m) * 64 + row * 8 + from % 8;
}

// DetEval for chess (definition shared with csrc/chess_tree.cuh): h = fold of mix64 over the bitboards and the
// side / castle / en-passant word; raw p(cell) = (1 + 3 bits of mix64(h ^ (2^32 + cell))) / 64; v on the 1/128 grid.
uint64_t det_hash(const Board& b) {
  uint64_t piece[6] = {0, 0, 0, 0, 0, 0}, color[2] = {0, 0};
  for (int s = 0; s < 64; ++s) {
    if (b.sq[s] == EMPTY) continue;
    piece[type_of(b.sq[s])] |= 1ull << s;
    color[color_of(b.sq[s])] |= 1ull << s;
  }
  uint64_t h = 0;
  for (int i = 0; i < 6; ++i) h = mix64(h ^ piece[i]);
  h = mix64(h ^ color[0]);
  h = mix64(h ^ color[1]);
  return mix64(h ^ ((uint64_t)b.side | ((uint64_t)b.castle << 8) | ((uint64_t)b.ep << 16)));
}

struct MNode {                         // mcts.rs:20-30
  Game state;
  long parent_id = -1;
  Move action_taken = 0xFFFF;
  float prior = 0.0f;
  std::vector<size_t> children_ids;
  uint32_t visit_count = 0;
  float value_sum = 0.0f;
  bool is_fully_expanded() const { return !children_ids.empty(); }
};

struct MTree {                         // mcts.rs:32-39
  std::vector<MNode> arena;
  float c = 2.0f;

  float get_ucb(size_t parent_id, size_t child_id) const {            // mcts.rs:91-100
    const MNode& parent = arena[parent_id];
    const MNode& child = arena[child_id];
    float q = 0.0f;
    if (child.visit_count != 0) q = (-child.value_sum / (float)child.visit_count + 1.0f) / 2.0f;
    return q + c * child.prior * std::sqrt((float)parent.visit_count) / (1.0f + (float)child.visit_count);
  }
  size_t select(size_t parent_id) const {                             // mcts.rs:102-114: Iterator::max_by keeps the LAST maximum
    const MNode& parent = arena[parent_id];
    size_t best = parent.children_ids[0];
    float best_s = get_ucb(parent_id, best);
    for (size_t i = 1; i < parent.children_ids.size(); ++i) {
      const float s = get_ucb(parent_id, parent.children_ids[i]);
      if (!(s < best_s)) { best = parent.children_ids[i]; best_s = s; }
    }
    return best;
  }
  void expand(size_t parent_id, const std::vector<float>& policy, uint64_t* children_created) {   // mcts.rs:116-143
    const Game parent_state = arena[parent_id].state;
    const std::vector<Move> actions = legal_moves(parent_state.b);    // get_valid_actions
    const size_t first = arena.size();
    for (size_t i = 0; i < actions.size(); ++i) arena[parent_id].children_ids.push_back(first + i);
    for (Move a : actions) {
      MNode ch;
      ch.state = parent_state;                                        // get_next_state, chess.rs:112-148
      ch.state.table.push_back(actions);
      ch.state.b = apply(parent_state.b, a);
      ch.parent_id = (long)parent_id;
      ch.action_taken = a;
      ch.prior = policy[policy_index(parent_state.b.side, a)];        // policy.get_prob(&action)
      arena.push_back(ch);
    }
    *children_created += actions.size();
  }
  void backprop(size_t node_id, float value) {                        // mcts.rs:145-159
    float sign = 1.0f;
    long id = (long)node_id;
    while (id >= 0) {
      arena[id].visit_count += 1;
      arena[id].value_sum += sign * value;
      sign *= -1.0f;
      id = arena[id].parent_id;
    }
  }
  void use_subtree(size_t new_root_id) {                              // mcts.rs:161-192
    std::vector<MNode> fresh;
    std::vector<size_t> queue_old{new_root_id};
    std::vector<long> queue_parent{-1};
    for (size_t head = 0; head < queue_old.size(); ++head) {
      MNode node = arena[queue_old[head]];
      const size_t id = fresh.size();
      node.parent_id = queue_parent[head];
      for (size_t child : node.children_ids) { queue_old.push_back(child); queue_parent.push_back((long)id); }
      node.children_ids.clear();
      if (node.parent_id >= 0) fresh[node.parent_id].children_ids.push_back(id);
      fresh.push_back(node);
    }
    arena.swap(fresh);
  }
};

}  // namespace

extern "C" {

typedef void (*orc_chess_eval_fn)(void* user, const float* encodings, uint32_t n, float* probs, float* values);

struct orc_chess_forest {
  std::vector<MTree> trees;
  uint64_t simulations = 0, evaluations = 0, terminal_leaves = 0, path_length_sum = 0, children_created = 0;
};

orc_chess_forest* orc_chess_forest_new(uint32_t n, float c) {
  orc_chess_forest* f = new orc_chess_forest();
  f->trees.resize(n);
  for (auto& t : f->trees) t.c = c;
  return f;
}
void orc_chess_forest_free(orc_chess_forest* f) { delete f; }

// Tree::with_root_state (mcts.rs:86-89)
void orc_chess_forest_reset(orc_chess_forest* f, uint32_t slot, const orc_chess* root) {
  MTree& t = f->trees[slot];
  t.arena.clear();
  MNode r;
  r.state = root->g;
  t.arena.push_back(r);
}

// Mcts::search (mcts.rs:196-332).  evaluator: SPB_EVAL_DET / SPB_EVAL_UNIFORM built in, SPB_EVAL_NET via fn (softmax
// output of the net over the 4,672 cells, NOT masked; the oracle applies mask_invalid_actions, chess.rs:251-271).
void orc_chess_forest_search(orc_chess_forest* f, uint32_t num_searches, int32_t evaluator, orc_chess_eval_fn fn, void* user) {
  for (uint32_t it = 0; it < num_searches; ++it) {                                 // :214
    std::vector<MTree*> to_expand;
    std::vector<size_t> node_ids;
    for (MTree& tree : f->trees) {                                                 // :236
      if (tree.arena.empty()) continue;
      size_t node = 0;
      while (tree.arena[node].is_fully_expanded()) { node = tree.select(node); f->path_length_sum++; }   // :239
      const int st = status(tree.arena[node].state);                               // get_value_and_terminated :243
      f->simulations++;
      if (st != SPB_STATUS_ONGOING) {
        tree.backprop(node, st == SPB_STATUS_WON ? 1.0f : 0.0f);                   // :246, chess.rs:172
        f->terminal_leaves++;
      } else {
        to_expand.push_back(&tree);
        node_ids.push_back(node);
      }
    }
    if (to_expand.empty()) continue;
    const size_t n = to_expand.size();
    std::vector<float> probs(n * POLICY), values(n, 0.0f);
    if (evaluator == SPB_EVAL_DET) {
      for (size_t i = 0; i < n; ++i) {
        const uint64_t h = det_hash(to_expand[i]->arena[node_ids[i]].state.b);
        for (int k = 0; k < POLICY; ++k) probs[i * POLICY + k] = (float)(1u + (uint32_t)(mix64(h ^ (0x100000000ull + (uint64_t)k)) & 7u)) * (1.0f / 64.0f);
        values[i] = ((float)((h >> 40) & 0xFFu) - 128.0f) * (1.0f / 128.0f);
      }
    } else if (evaluator == SPB_EVAL_UNIFORM) {
      for (auto& p : probs) p = 1.0f;
    } else {
      std::vector<float> enc(n * 19 * 64);
      for (size_t i = 0; i < n; ++i) {
        orc_chess tmp;
        tmp.g = to_expand[i]->arena[node_ids[i]].state;
        orc_chess_encode(&tmp, &enc[i * 19 * 64]);                                  // model/mod.rs:41-44
      }
      fn(user, enc.data(), (uint32_t)n, probs.data(), values.data());              // :60-67, :95
    }
    f->evaluations += n;
    for (size_t i = 0; i < n; ++i) {                                               // :278-284
      MTree* tree = to_expand[i];
      const Game& st = tree->arena[node_ids[i]].state;
      // mask_invalid_actions, chess.rs:251-271: probs * mask, divided by the ndarray sum of the masked array
      std::vector<float> masked(POLICY, 0.0f);
      for (Move m : legal_moves(st.b)) { const int k = policy_index(st.b.side, m); masked[k] = probs[i * POLICY + k] * 1.0f; }
      const float sum = ndarray_sum(masked.data(), POLICY);
      for (auto& x : masked) x = x / sum;
      tree->expand(node_ids[i], masked, &f->children_created);
      tree->backprop(node_ids[i], values[i]);
    }
  }
}

int32_t orc_chess_forest_root_children(const orc_chess_forest* f, uint32_t slot, uint16_t* moves, uint32_t* counts, uint32_t* ids) {
  const MTree& t = f->trees[slot];
  const auto& ch = t.arena[0].children_ids;
  for (size_t i = 0; i < ch.size(); ++i) {
    if (moves) moves[i] = t.arena[ch[i]].action_taken;
    if (counts) counts[i] = t.arena[ch[i]].visit_count;
    if (ids) ids[i] = (uint32_t)ch[i];
  }
  return (int32_t)ch.size();
}
uint32_t orc_chess_forest_arena_len(const orc_chess_forest* f, uint32_t slot) { return (uint32_t)f->trees[slot].arena.size(); }
int32_t orc_chess_forest_node(const orc_chess_forest* f, uint32_t slot, uint32_t id, uint32_t* n, float* w, float* p, uint32_t* first_child,
                              uint32_t* n_children, uint16_t* move) {
  const MTree& t = f->trees[slot];
  if (id >= t.arena.size()) return SPB_ERR_ARG;
  const MNode& nd = t.arena[id];
  *n = nd.visit_count; *w = nd.value_sum; *p = nd.prior;
  *first_child = nd.children_ids.empty() ? 0u : (uint32_t)nd.children_ids[0];
  *n_children = (uint32_t)nd.children_ids.size();
  *move = nd.action_taken;
  return SPB_OK;
}
void orc_chess_forest_use_subtree(orc_chess_forest* f, uint32_t slot, uint32_t id) { f->trees[slot].use_subtree(id); }
orc_chess* orc_chess_forest_state(const orc_chess_forest* f, uint32_t slot, uint32_t id) {
  orc_chess* c = new orc_chess();
  c->g = f->trees[slot].arena[id].state;
  return c;
}
void orc_chess_forest_counters(const orc_chess_forest* f, uint64_t* out) {
  out[0] = f->simulations; out[1] = f->evaluations; out[2] = f->terminal_leaves; out[3] = f->path_length_sum; out[4] = f->children_created;
}
uint64_t orc_chess_det_hash(const orc_chess* c) { return det_hash(c->g.b); }

}  // extern "C"
