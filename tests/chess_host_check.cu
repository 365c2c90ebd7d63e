// Test harness (NOT product): compiles the __host__ __device__ rule code of csrc/chess.cuh / chess_tree.cuh for the HOST so
// that tests/test_chess.py can compare it with the oracle on a machine without a GPU.  The product never runs these
// functions on the host (spb_chess_* launch kernels); this only shortens the loop for the shared rules code.
#include "../self-play-ai_b200/csrc/chess_tree.cuh"

using namespace spb::chess;

static unsigned long long perft_rec(const Pos& p, int depth) {
  Move mv[MAX_MOVES];
  const int n = legal_moves(p, mv);
  if (depth <= 1) return depth == 1 ? (unsigned long long)n : 1ull;
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i) total += perft_rec(apply_move(p, mv[i]), depth - 1);
  return total;
}

extern "C" {
int host_legal(const spb_chess_state* p, uint16_t* out) { return legal_moves(*p, out); }
void host_apply(const spb_chess_state* p, uint16_t mv, spb_chess_state* out) { *out = apply_move(*p, mv); }
unsigned long long host_perft(const spb_chess_state* p, int depth) { return perft_rec(*p, depth); }
unsigned long long host_list_hash(const uint16_t* mv, int n) { return move_list_hash(mv, n); }
float host_encode(const spb_chess_state* p, uint32_t reps, int plane, int row, int col) { return encode_plane(*p, reps, plane, row, col); }
unsigned long long host_det_hash(const spb_chess_state* p) { return det_hash(*p); }
float host_det_raw_prob(unsigned long long h, int idx) { return det_raw_prob(h, idx); }
float host_det_value(unsigned long long h) { return det_value(h); }
int host_in_check(const spb_chess_state* p) { return in_check(*p, p->side) ? 1 : 0; }
}
