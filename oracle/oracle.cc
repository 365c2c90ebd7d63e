// oracle.cc — CPU restatement of the reference's self-play hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT (see oracle.h).  PARITY UNPINNED BY THE REFERENCE.
//
// Follows, line by line:
//   src/mcts.rs:91-159 (get_ucb / select / expand / backprop), :161-192 (use_subtree),
//   :214-331 (lock-step search loop and result assembly),
//   src/game/connect_four.rs:127-283, src/game/tictactoe.rs:127-241,
//   src/model/mod.rs:36-98 (predict plumbing), src/main.rs:108-112 (greedy last-max rule).
// It keeps the reference's *structure* on purpose (AoS arena, a full state copy per node,
// array boards, a serial loop over trees, one evaluator batch per simulation step) so that
// it can also serve as the CPU baseline.  All arithmetic is f32 with -ffp-contract=off.
#include "oracle.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

namespace {

enum : uint8_t { ONGOING = SPB_STATUS_ONGOING, TIED = SPB_STATUS_TIED, WON = SPB_STATUS_WON };
constexpr int8_t NONE = -1;  // Piece(None)

// ---- splitmix64 + DetEval -------------------------------------------------
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ndarray 0.15.6 `sum()` on a contiguous f32 slice: numeric_util::unrolled_fold — eight partial
// sums over stride-8 lanes, combined as ((((0+(p0+p4))+(p1+p5))+(p2+p6))+(p3+p7)), then the
// (<8) tail added sequentially.  Call sites: connect_four.rs:97,276; tictactoe.rs:97,233.
inline float ndarray_sum(const float* xs, size_t n) {
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

// ---- Connect4 (src/game/connect_four.rs) ---------------------------------------------------
struct C4State {
  static constexpr int GAME = SPB_GAME_CONNECT4;
  static constexpr int A = 7;
  static constexpr int ROWS = 6, COLS = 7;
  int8_t board[6][7];  // connect_four.rs:18  Board([[Piece;7];6]); row 0 = bottom (:52-65)
  uint8_t current_player = 0;       // :23
  uint8_t num_actions_played = 0;   // :24
  uint8_t status = ONGOING;         // :25

  C4State() { std::memset(board, NONE, sizeof board); }

  bool operator==(const C4State& o) const {
    return std::memcmp(board, o.board, sizeof board) == 0 && current_player == o.current_player &&
           num_actions_played == o.num_actions_played && status == o.status;
  }

  // connect_four.rs:128-136
  int get_next_row_idx(int col) const {
    for (int i = 0; i < ROWS; ++i)
      if (board[i][col] == NONE) return i;
    return -1;
  }

  // connect_four.rs:140-179.  NOTE: only the (row+i, col+i) diagonal is checked (:163-176).
  int8_t get_winner(int latest_row, int latest_col) const {
    const int8_t* row = board[latest_row];
    for (int i = 0; i <= COLS - 4; ++i)  // :143
      if (row[i] != NONE && row[i] == row[i + 1] && row[i] == row[i + 2] && row[i] == row[i + 3]) return row[i];
    for (int i = 0; i <= ROWS - 4; ++i)  // :153
      if (board[i][latest_col] != NONE && board[i][latest_col] == board[i + 1][latest_col] &&
          board[i][latest_col] == board[i + 2][latest_col] && board[i][latest_col] == board[i + 3][latest_col])
        return board[i][latest_col];
    // :164-165
    int start_offset = std::max(-4, -std::min(latest_col, latest_row));
    int end_offset = std::min(0, std::min(COLS - (latest_col + 4), ROWS - (latest_row + 4)));
    for (int i = start_offset; i <= end_offset; ++i) {  // :166
      int r = latest_row + i, c = latest_col + i;
      if (board[r][c] != NONE && board[r][c] == board[r + 1][c + 1] && board[r][c] == board[r + 2][c + 2] &&
          board[r][c] == board[r + 3][c + 3])
        return board[r][c];
    }
    return NONE;
  }

  // connect_four.rs:190-211.  Returns false for Err(..).
  bool get_next_state(int action, C4State* out) const {
    if (status != ONGOING) return false;          // :209
    if (action < 0 || action >= COLS) return false;
    int row_idx = get_next_row_idx(action);
    if (row_idx < 0) return false;                // :193
    C4State next = *this;                         // :195
    next.board[row_idx][action] = (int8_t)current_player;
    next.current_player = current_player ^ 1;     // :197
    next.num_actions_played += 1;
    if (next.get_winner(row_idx, action) != NONE) next.status = WON;   // :200
    else if (next.num_actions_played == 6 * 7) next.status = TIED;     // :202
    *out = next;
    return true;
  }

  // connect_four.rs:213-225
  int get_valid_actions(int* actions) const {
    int n = 0;
    if (status == ONGOING)
      for (int col = 0; col < 7; ++col)
        if (board[ROWS - 1][col] == NONE) actions[n++] = col;
    return n;
  }

  // connect_four.rs:231-240
  void get_value_and_terminated(float* v, bool* term) const {
    if (status == WON) { *v = -1.0f; *term = true; }
    else if (status == TIED) { *v = 0.0f; *term = true; }
    else { *v = 0.0f; *term = false; }
  }

  // connect_four.rs:242-259 — [plane][row][col]
  void get_encoding(float* e) const {
    std::memset(e, 0, sizeof(float) * 3 * ROWS * COLS);
    for (int row = 0; row < ROWS; ++row)
      for (int col = 0; col < COLS; ++col) {
        int8_t p = board[row][col];
        int plane = (p == NONE) ? 2 : (p == (int8_t)current_player ? 0 : 1);
        e[(plane * ROWS + row) * COLS + col] = 1.0f;
      }
  }

  // connect_four.rs:261-279
  void mask_invalid_actions(const float* policy, float* out) const {
    int acts[7];
    int n = get_valid_actions(acts);
    float mask[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) mask[acts[i]] = 1.0f;
    float masked[7];
    for (int i = 0; i < 7; ++i) masked[i] = policy[i] * mask[i];   // :275
    float s = ndarray_sum(masked, 7);                              // :276
    for (int i = 0; i < 7; ++i) out[i] = masked[i] / s;
  }

  static int bit(int row, int col) { return col * 7 + row; }

  void to_abi(spb_state* s) const {
    std::memset(s, 0, sizeof *s);
    for (int r = 0; r < ROWS; ++r)
      for (int c = 0; c < COLS; ++c)
        if (board[r][c] != NONE) s->stones[board[r][c]] |= 1ull << bit(r, c);
    s->current_player = current_player;
    s->num_actions_played = num_actions_played;
    s->status = status;
  }
  static C4State from_abi(const spb_state& s) {
    C4State st;
    for (int r = 0; r < ROWS; ++r)
      for (int c = 0; c < COLS; ++c) {
        if (s.stones[0] >> bit(r, c) & 1) st.board[r][c] = 0;
        else if (s.stones[1] >> bit(r, c) & 1) st.board[r][c] = 1;
      }
    st.current_player = s.current_player;
    st.num_actions_played = s.num_actions_played;
    st.status = s.status;
    return st;
  }
};

// ---- Tic-tac-toe (src/game/tictactoe.rs) ---------------------------------------------------
struct TttState {
  static constexpr int GAME = SPB_GAME_TICTACTOE;
  static constexpr int A = 9;
  static constexpr int ROWS = 3, COLS = 3;
  int8_t board[3][3];  // tictactoe.rs:18
  uint8_t current_player = 0;
  uint8_t num_actions_played = 0;
  uint8_t status = ONGOING;

  TttState() { std::memset(board, NONE, sizeof board); }
  bool operator==(const TttState& o) const {
    return std::memcmp(board, o.board, sizeof board) == 0 && current_player == o.current_player &&
           num_actions_played == o.num_actions_played && status == o.status;
  }

  // tictactoe.rs:135-167; action = row*3+col (Action{row,col}, :28-32; flat index :113,:123)
  bool get_next_state(int action, TttState* out) const {
    if (status != ONGOING) return false;   // :165
    if (action < 0 || action >= 9) return false;
    int ar = action / 3, ac = action % 3;
    if (board[ar][ac] != NONE) return false;  // :138
    TttState next = *this;
    next.board[ar][ac] = (int8_t)current_player;
    next.current_player = current_player ^ 1;
    next.num_actions_played += 1;
    const int8_t* row = next.board[ar];
    bool is_row_win = row[0] == row[1] && row[1] == row[2];                                   // :146
    bool is_col_win = next.board[0][ac] == next.board[1][ac] && next.board[1][ac] == next.board[2][ac];  // :148-150
    bool is_nw_se = ar == ac && next.board[0][0] == next.board[1][1] && next.board[1][1] == next.board[2][2];  // :152
    int diff = ar > ac ? ar - ac : ac - ar;
    bool is_ne_sw = ((ar == 1 && ac == 1) || diff == 2) && next.board[0][2] == next.board[1][1] &&
                    next.board[1][1] == next.board[2][0];                                     // :154
    if (is_row_win || is_col_win || is_nw_se || is_ne_sw) next.status = WON;                  // :157
    else if (next.num_actions_played == 9) next.status = TIED;                                // :159
    *out = next;
    return true;
  }

  // tictactoe.rs:169-182 (row-major)
  int get_valid_actions(int* actions) const {
    int n = 0;
    if (status == ONGOING)
      for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
          if (board[row][col] == NONE) actions[n++] = row * 3 + col;
    return n;
  }

  void get_value_and_terminated(float* v, bool* term) const {  // :188-197
    if (status == WON) { *v = -1.0f; *term = true; }
    else if (status == TIED) { *v = 0.0f; *term = true; }
    else { *v = 0.0f; *term = false; }
  }

  void get_encoding(float* e) const {  // :199-216
    std::memset(e, 0, sizeof(float) * 27);
    for (int row = 0; row < 3; ++row)
      for (int col = 0; col < 3; ++col) {
        int8_t p = board[row][col];
        int plane = (p == NONE) ? 2 : (p == (int8_t)current_player ? 0 : 1);
        e[(plane * 3 + row) * 3 + col] = 1.0f;
      }
  }

  void mask_invalid_actions(const float* policy, float* out) const {  // :218-236
    int acts[9];
    int n = get_valid_actions(acts);
    float mask[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) mask[acts[i]] = 1.0f;
    float masked[9];
    for (int i = 0; i < 9; ++i) masked[i] = policy[i] * mask[i];
    float s = ndarray_sum(masked, 9);   // :233 (Array2 (3,3) contiguous -> slice fold over 9)
    for (int i = 0; i < 9; ++i) out[i] = masked[i] / s;
  }

  static int bit(int row, int col) { return row * 3 + col; }
  void to_abi(spb_state* s) const {
    std::memset(s, 0, sizeof *s);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        if (board[r][c] != NONE) s->stones[board[r][c]] |= 1ull << bit(r, c);
    s->current_player = current_player;
    s->num_actions_played = num_actions_played;
    s->status = status;
  }
  static TttState from_abi(const spb_state& s) {
    TttState st;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        if (s.stones[0] >> bit(r, c) & 1) st.board[r][c] = 0;
        else if (s.stones[1] >> bit(r, c) & 1) st.board[r][c] = 1;
      }
    st.current_player = s.current_player;
    st.num_actions_played = s.num_actions_played;
    st.status = s.status;
    return st;
  }
};

template <class S>
uint64_t det_hash(const S& st) {
  spb_state a;
  st.to_abi(&a);
  uint64_t mine = a.stones[st.current_player], opp = a.stones[st.current_player ^ 1];
  return splitmix64(mine ^ splitmix64(opp));
}
template <class S>
void det_eval(const S& st, float* probs, float* value) {
  uint64_t h = det_hash(st);
  for (int a = 0; a < S::A; ++a) probs[a] = (float)(1 + ((h >> (4 * a)) & 7)) / 64.0f;
  *value = ((float)((h >> 40) & 0xFF) - 128.0f) / 128.0f;
}

// ---- mcts.rs --------------------------------------------------------------------------------
template <class S>
struct Node {                       // mcts.rs:20-30
  S state;
  size_t id = 0;
  long parent_id = -1;              // Option<usize>
  int action_taken = -1;            // Option<Action>
  float prior = 0.0f;               // Option<f32>
  bool has_prior = false;
  std::vector<size_t> children_ids;
  uint32_t visit_count = 0;
  float value_sum = 0.0f;
  bool is_fully_expanded() const { return !children_ids.empty(); }   // :62-64
};

struct Counters {
  uint64_t simulations = 0, evaluations = 0, terminal_leaves = 0, path_length_sum = 0, children_created = 0;
};

template <class S>
struct Tree {                       // mcts.rs:32-39
  float c = 2.0f;                   // args.c of the TREE (mcts.rs:99; Args::default :49)
  std::vector<Node<S>> arena;
  long node_id_to_expand = -1;

  Tree() { arena.emplace_back(); }                        // Tree::default :67-77
  explicit Tree(const S& s) { arena.emplace_back(); arena[0].state = s; }   // with_root_state :86-89

  // mcts.rs:91-100
  float get_ucb(size_t parent_id, size_t child_id) const {
    const Node<S>& parent = arena[parent_id];
    const Node<S>& child = arena[child_id];
    float q;
    if (child.visit_count == 0) q = 0.0f;
    else q = (-child.value_sum / (float)child.visit_count + 1.0f) / 2.0f;
    float u = c * child.prior;
    u = u * std::sqrt((float)parent.visit_count);
    u = u / (1.0f + (float)child.visit_count);
    return q + u;
  }

  // mcts.rs:102-114 — Iterator::max_by returns the LAST maximal element.
  size_t select(size_t parent_id) const {
    const Node<S>& parent = arena[parent_id];
    size_t best = parent.children_ids[0];
    for (size_t i = 1; i < parent.children_ids.size(); ++i) {
      size_t cand = parent.children_ids[i];
      float a = get_ucb(parent_id, best), b = get_ucb(parent_id, cand);
      if (std::isnan(a) || std::isnan(b)) { std::fprintf(stderr, "oracle: NaN ucb (reference would panic)\n"); std::abort(); }
      if (!(a > b)) best = cand;   // Ordering::Greater keeps the accumulator, otherwise take the new one
    }
    return best;
  }

  // mcts.rs:116-143
  void expand(size_t parent_id, const float* policy, Counters& ctr) {
    size_t arena_len = arena.size();
    int actions[S::A];
    S parent_state = arena[parent_id].state;      // :124
    int n = parent_state.get_valid_actions(actions);
    for (int i = 0; i < n; ++i) arena[parent_id].children_ids.push_back(arena_len + i);   // :121-122
    size_t new_id = arena_len;
    for (int i = 0; i < n; ++i) {
      Node<S> child;
      bool ok = parent_state.get_next_state(actions[i], &child.state);   // :129 unwrap
      if (!ok) { std::fprintf(stderr, "oracle: get_next_state failed in expand\n"); std::abort(); }
      child.id = new_id;
      child.parent_id = (long)parent_id;
      child.action_taken = actions[i];
      child.prior = policy[actions[i]];           // :128 policy.get_prob(&action)
      child.has_prior = true;
      arena.push_back(std::move(child));
      ++new_id;
    }
    ctr.children_created += n;
  }

  // mcts.rs:145-159
  void backprop(size_t node_id, float value) {
    Node<S>* node = &arena[node_id];
    float sign = 1.0f;
    node->visit_count += 1;
    node->value_sum += sign * value;
    sign *= -1.0f;
    while (node->parent_id >= 0) {
      node = &arena[node->parent_id];
      node->visit_count += 1;
      node->value_sum += sign * value;
      sign *= -1.0f;
    }
  }

  // mcts.rs:161-192
  void use_subtree(size_t new_root_id) {
    std::vector<Node<S>>& old_arena = arena;
    std::vector<Node<S>> new_arena;
    Node<S> new_root = old_arena[new_root_id];
    new_root.parent_id = -1;
    std::deque<Node<S>> nodes_to_add;
    nodes_to_add.push_back(std::move(new_root));
    size_t next_id = 0;
    while (!nodes_to_add.empty()) {
      Node<S> node = std::move(nodes_to_add.front());
      nodes_to_add.pop_front();
      node.id = next_id;
      for (size_t child_id : node.children_ids) {
        Node<S> child = old_arena[child_id];
        child.parent_id = (long)node.id;
        nodes_to_add.push_back(std::move(child));
      }
      node.children_ids.clear();
      if (node.parent_id >= 0) new_arena[node.parent_id].children_ids.push_back(node.id);
      new_arena.push_back(std::move(node));
      ++next_id;
    }
    arena = std::move(new_arena);
  }
};

template <class S>
struct Evaluator {
  int kind;
  orc_eval_fn fn;
  void* user;
  // Model::predict model/mod.rs:36-98: encode -> forward+softmax (or DetEval/uniform) -> per-state mask.
  void predict(const std::vector<const S*>& states, std::vector<float>& policies, std::vector<float>& values) const {
    size_t n = states.size();
    constexpr int A = S::A;
    std::vector<float> probs(n * A);
    values.assign(n, 0.0f);
    if (kind == SPB_EVAL_DET) {
      for (size_t i = 0; i < n; ++i) det_eval(*states[i], &probs[i * A], &values[i]);
    } else if (kind == SPB_EVAL_UNIFORM) {
      for (size_t i = 0; i < n * A; ++i) probs[i] = 1.0f;
    } else {
      constexpr int E = 3 * S::ROWS * S::COLS;
      std::vector<float> enc(n * E);
      for (size_t i = 0; i < n; ++i) states[i]->get_encoding(&enc[i * E]);    // :41-44
      fn(user, enc.data(), (uint32_t)n, probs.data(), values.data());         // :60-67, :95
    }
    policies.resize(n * A);
    for (size_t i = 0; i < n; ++i) states[i]->mask_invalid_actions(&probs[i * A], &policies[i * A]);   // :86-93
  }
};

// Mcts::search mcts.rs:196-332 (result assembly is done by the accessors below)
template <class S>
void search(std::vector<Tree<S>*>& trees, uint32_t num_searches, const Evaluator<S>& ev, Counters& ctr) {
  constexpr int A = S::A;
  std::vector<Tree<S>*> trees_to_expand;
  std::vector<const S*> states;
  std::vector<float> policies, values;
  for (uint32_t it = 0; it < num_searches; ++it) {             // :214
    trees_to_expand.clear();
    for (Tree<S>* tree : trees) {                              // :236
      size_t node = 0;
      while (tree->arena[node].is_fully_expanded()) {          // :239
        node = tree->select(node);
        ctr.path_length_sum++;
      }
      float value; bool is_terminal;
      tree->arena[node].state.get_value_and_terminated(&value, &is_terminal);   // :243
      ctr.simulations++;
      if (is_terminal) {
        tree->backprop(node, value);                           // :246
        tree->node_id_to_expand = -1;
        ctr.terminal_leaves++;
      } else {
        tree->node_id_to_expand = (long)node;                  // :249
        trees_to_expand.push_back(tree);
      }
    }
    if (!trees_to_expand.empty()) {                            // :254
      states.clear();
      for (Tree<S>* tree : trees_to_expand) states.push_back(&tree->arena[tree->node_id_to_expand].state);
      ev.predict(states, policies, values);                    // :268
      ctr.evaluations += states.size();
      for (size_t i = 0; i < trees_to_expand.size(); ++i) {    // :278-284
        Tree<S>* tree = trees_to_expand[i];
        size_t node_id = (size_t)tree->node_id_to_expand;
        tree->expand(node_id, &policies[i * A], ctr);
        tree->backprop(node_id, values[i]);
      }
    }
  }
}

// ---- EXTENSION (not in the reference): K in-flight leaves per tree with virtual loss --------------------
// BASELINE.json config 4 / SURVEY.md §8(f)-3.  The reference runs exactly one leaf per tree per step
// (mcts.rs:236-252); with K = 1 this function is never used and the reference algorithm above is.  Definition
// (the device kernels implement exactly the same operation order):
//   a step performs k_this = min(K, remaining) simulations per tree.  Simulation k descends with PUCT on the
//   current statistics.  Terminal leaf: real backprop at once.  Otherwise every node of the path gets a
//   virtual loss (visit_count += 1, value_sum += 1.0) and the leaf is queued; if the same leaf was already
//   queued in this step the simulation is a duplicate of that entry (it still takes its virtual loss).
//   After the batch evaluation, entries are finished in selection order: a first occurrence expands the leaf;
//   every entry then rewrites each path node as value_sum = (value_sum - 1.0) + sign*v (the visit stays).
template <class S>
void search_vl(std::vector<Tree<S>*>& trees, uint32_t num_searches, uint32_t K, const Evaluator<S>& ev, Counters& ctr) {
  constexpr int A = S::A;
  struct Entry { size_t leaf; int dup_of; std::vector<size_t> path; };
  std::vector<std::vector<Entry>> pending(trees.size());
  std::vector<const S*> states;
  std::vector<std::pair<size_t, size_t>> owners;   // (tree index, entry index) per evaluated state
  std::vector<float> policies, values;
  uint32_t done = 0;
  while (done < num_searches) {
    const uint32_t k_this = std::min(K, num_searches - done);
    states.clear(); owners.clear();
    for (size_t ti = 0; ti < trees.size(); ++ti) {
      Tree<S>* tree = trees[ti];
      auto& pend = pending[ti];
      pend.clear();
      for (uint32_t k = 0; k < k_this; ++k) {
        std::vector<size_t> path{0};
        size_t node = 0;
        while (tree->arena[node].is_fully_expanded()) {
          node = tree->select(node);
          path.push_back(node);
          ctr.path_length_sum++;
        }
        ctr.simulations++;
        float value; bool is_terminal;
        tree->arena[node].state.get_value_and_terminated(&value, &is_terminal);
        if (is_terminal) {
          tree->backprop(node, value);
          ctr.terminal_leaves++;
          continue;
        }
        int dup = -1;
        for (size_t j = 0; j < pend.size(); ++j)
          if (pend[j].dup_of < 0 && pend[j].leaf == node) { dup = (int)j; break; }
        for (size_t id : path) { tree->arena[id].visit_count += 1; tree->arena[id].value_sum += 1.0f; }   // virtual loss
        pend.push_back(Entry{node, dup, path});
        if (dup < 0) { states.push_back(&tree->arena[node].state); owners.emplace_back(ti, pend.size() - 1); }
      }
    }
    if (!states.empty()) {
      ev.predict(states, policies, values);
      ctr.evaluations += states.size();
      std::vector<std::vector<float>> val_of(trees.size());
      std::vector<std::vector<const float*>> pol_of(trees.size());
      for (size_t ti = 0; ti < trees.size(); ++ti) { val_of[ti].assign(pending[ti].size(), 0.0f); pol_of[ti].assign(pending[ti].size(), nullptr); }
      for (size_t i = 0; i < owners.size(); ++i) { val_of[owners[i].first][owners[i].second] = values[i]; pol_of[owners[i].first][owners[i].second] = &policies[i * A]; }
      for (size_t ti = 0; ti < trees.size(); ++ti) {
        Tree<S>* tree = trees[ti];
        auto& pend = pending[ti];
        for (size_t e = 0; e < pend.size(); ++e) {
          const size_t src = pend[e].dup_of < 0 ? e : (size_t)pend[e].dup_of;
          if (pend[e].dup_of < 0) tree->expand(pend[e].leaf, pol_of[ti][e], ctr);
          const float v = val_of[ti][src];
          const auto& path = pend[e].path;
          float sign = 1.0f;
          for (size_t d = path.size(); d-- > 0;) {          // leaf first (+v), then parents with alternating sign
            Node<S>& nd = tree->arena[path[d]];
            nd.value_sum = (nd.value_sum - 1.0f) + sign * v;
            sign *= -1.0f;
          }
        }
      }
    }
    done += k_this;
  }
}

struct ForestBase {
  std::string err;
  int game;
  virtual ~ForestBase() {}
  virtual int32_t reset(const uint32_t* slots, uint32_t n, const spb_state* roots) = 0;
  virtual int32_t do_search(uint32_t s, int32_t kind, orc_eval_fn fn, void* user) = 0;
  virtual int32_t root_children(uint32_t slot, uint8_t*, uint32_t*, uint32_t*, uint32_t*) = 0;
  virtual int32_t root_policy(uint32_t slot, float*) = 0;
  virtual int32_t use_subtree(uint32_t slot, uint32_t node) = 0;
  virtual int32_t get_state(uint32_t slot, uint32_t node, spb_state*) = 0;
  virtual int32_t arena_len(uint32_t slot, uint32_t*) = 0;
  virtual int32_t node_stats(uint32_t, uint32_t, uint32_t*, float*, float*, uint32_t*, uint32_t*) = 0;
  virtual void counters(spb_counters*) = 0;
  virtual void set_leaves_per_tree(uint32_t k) = 0;
};

template <class S>
struct Forest : ForestBase {
  std::vector<Tree<S>> trees;
  float c;
  uint32_t K = 1;
  Counters ctr;
  Forest(uint32_t n, float c_) : trees(n), c(c_) { game = S::GAME; for (auto& t : trees) t.c = c; }
  bool ok(uint32_t slot) { if (slot >= trees.size()) { err = "slot out of range"; return false; } return true; }
  int32_t reset(const uint32_t* slots, uint32_t n, const spb_state* roots) override {
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t s = slots ? slots[i] : i;
      if (!ok(s)) return SPB_ERR_ARG;
      trees[s] = roots ? Tree<S>(S::from_abi(roots[i])) : Tree<S>();
      trees[s].c = c;
    }
    return SPB_OK;
  }
  int32_t do_search(uint32_t s, int32_t kind, orc_eval_fn fn, void* user) override {
    if (kind == SPB_EVAL_NET && !fn) { err = "SPB_EVAL_NET needs a callback"; return SPB_ERR_ARG; }
    std::vector<Tree<S>*> ptrs;
    for (auto& t : trees) ptrs.push_back(&t);
    Evaluator<S> ev{kind, fn, user};
    if (K <= 1) search(ptrs, s, ev, ctr); else search_vl(ptrs, s, K, ev, ctr);
    return SPB_OK;
  }
  // mcts.rs:310-331
  int32_t root_children(uint32_t slot, uint8_t* actions, uint32_t* counts, uint32_t* ids, uint32_t* n) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    const Tree<S>& t = trees[slot];
    const auto& ch = t.arena[0].children_ids;
    for (size_t i = 0; i < ch.size(); ++i) {
      const Node<S>& cn = t.arena[ch[i]];
      if (actions) actions[i] = (uint8_t)cn.action_taken;
      if (counts) counts[i] = cn.visit_count;
      if (ids) ids[i] = (uint32_t)ch[i];
    }
    if (n) *n = (uint32_t)ch.size();
    return SPB_OK;
  }
  int32_t root_policy(uint32_t slot, float* policy) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    const Tree<S>& t = trees[slot];
    float p[S::A];
    for (int i = 0; i < S::A; ++i) p[i] = 0.0f;                               // get_zero_policy :315
    for (size_t cid : t.arena[0].children_ids) p[t.arena[cid].action_taken] = (float)t.arena[cid].visit_count;  // :322-324
    float s = ndarray_sum(p, S::A);                                           // normalize :328
    for (int i = 0; i < S::A; ++i) policy[i] = p[i] / s;
    return SPB_OK;
  }
  int32_t use_subtree(uint32_t slot, uint32_t node) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    if (node >= trees[slot].arena.size()) { err = "node id out of range"; return SPB_ERR_ARG; }
    trees[slot].node_id_to_expand = -1;
    trees[slot].use_subtree(node);
    return SPB_OK;
  }
  int32_t get_state(uint32_t slot, uint32_t node, spb_state* out) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    if (node >= trees[slot].arena.size()) { err = "node id out of range"; return SPB_ERR_ARG; }
    trees[slot].arena[node].state.to_abi(out);
    return SPB_OK;
  }
  int32_t arena_len(uint32_t slot, uint32_t* out) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    *out = (uint32_t)trees[slot].arena.size();
    return SPB_OK;
  }
  int32_t node_stats(uint32_t slot, uint32_t node, uint32_t* n, float* w, float* p, uint32_t* fc, uint32_t* nc) override {
    if (!ok(slot)) return SPB_ERR_ARG;
    if (node >= trees[slot].arena.size()) { err = "node id out of range"; return SPB_ERR_ARG; }
    const Node<S>& nd = trees[slot].arena[node];
    if (n) *n = nd.visit_count;
    if (w) *w = nd.value_sum;
    if (p) *p = nd.prior;
    if (fc) *fc = nd.children_ids.empty() ? 0u : (uint32_t)nd.children_ids[0];
    if (nc) *nc = (uint32_t)nd.children_ids.size();
    return SPB_OK;
  }
  void set_leaves_per_tree(uint32_t k) override { K = k; }
  void counters(spb_counters* out) override {
    std::memset(out, 0, sizeof *out);
    out->simulations = ctr.simulations;
    out->evaluations = ctr.evaluations;
    out->terminal_leaves = ctr.terminal_leaves;
    out->path_length_sum = ctr.path_length_sum;
    out->children_created = ctr.children_created;
    for (auto& t : trees) out->nodes_live += t.arena.size();
  }
};

// main.rs:106-114 greedy: max_by(total_cmp) over (child id, count) -> last max; use_subtree.
template <class S>
int32_t greedy_game(const spb_state* root, float c, uint32_t num_searches, int32_t kind, orc_eval_fn fn, void* user,
                    uint8_t* actions, uint32_t* arena_sizes, uint32_t* last_counts, uint32_t* n_last, uint8_t* final_status) {
  Tree<S> tree = root ? Tree<S>(S::from_abi(*root)) : Tree<S>();
  tree.c = c;
  Counters ctr;
  Evaluator<S> ev{kind, fn, user};
  int ply = 0;
  for (;;) {
    std::vector<Tree<S>*> v{&tree};
    search(v, num_searches, ev, ctr);
    if (arena_sizes) arena_sizes[ply] = (uint32_t)tree.arena.size();
    const auto& ch = tree.arena[0].children_ids;
    if (ch.empty()) break;
    size_t best = 0;
    for (size_t i = 1; i < ch.size(); ++i)
      if (!(tree.arena[ch[best]].visit_count > tree.arena[ch[i]].visit_count)) best = i;
    if (last_counts) for (size_t i = 0; i < ch.size(); ++i) last_counts[i] = tree.arena[ch[i]].visit_count;
    if (n_last) *n_last = (uint32_t)ch.size();
    size_t best_id = ch[best];
    tree.use_subtree(best_id);
    actions[ply++] = (uint8_t)tree.arena[0].action_taken;
    if (tree.arena[0].state.status != ONGOING) break;
    if (ply >= 64) break;
  }
  if (final_status) *final_status = tree.arena[0].state.status;
  return ply;
}

template <class S>
uint64_t baseline_run(const spb_state* roots, uint32_t threads, uint32_t gpt, float c, uint32_t num_searches,
                      int32_t kind, orc_eval_fn fn, void* user, double* seconds) {
  std::vector<Forest<S>*> forests;
  for (uint32_t t = 0; t < threads; ++t) {
    auto* f = new Forest<S>(gpt, c);
    if (roots) f->reset(nullptr, gpt, roots + (size_t)t * gpt);
    forests.push_back(f);
  }
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; ++t)
    th.emplace_back([&, t] { forests[t]->do_search(num_searches, kind, fn, user); });
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  uint64_t sims = 0;
  for (auto* f : forests) { sims += f->ctr.simulations; delete f; }
  return sims;
}

}  // namespace

// ---- C interface ----------------------------------------------------------------------------
struct orc_forest { ForestBase* impl; };

extern "C" {

orc_forest* orc_create(int32_t game, uint32_t num_trees, float c) {
  ForestBase* impl = nullptr;
  if (game == SPB_GAME_CONNECT4) impl = new Forest<C4State>(num_trees, c);
  else if (game == SPB_GAME_TICTACTOE) impl = new Forest<TttState>(num_trees, c);
  else return nullptr;
  return new orc_forest{impl};
}
void orc_destroy(orc_forest* f) { if (f) { delete f->impl; delete f; } }
const char* orc_last_error(const orc_forest* f) { return f ? f->impl->err.c_str() : "null forest"; }
int32_t orc_reset(orc_forest* f, const uint32_t* slots, uint32_t n, const spb_state* roots) { return f->impl->reset(slots, n, roots); }
int32_t orc_search(orc_forest* f, uint32_t s, int32_t kind, orc_eval_fn fn, void* user) { return f->impl->do_search(s, kind, fn, user); }
int32_t orc_root_children(orc_forest* f, uint32_t slot, uint8_t* a, uint32_t* c, uint32_t* ids, uint32_t* n) { return f->impl->root_children(slot, a, c, ids, n); }
int32_t orc_root_policy(orc_forest* f, uint32_t slot, float* p) { return f->impl->root_policy(slot, p); }
int32_t orc_use_subtree(orc_forest* f, uint32_t slot, uint32_t node) { return f->impl->use_subtree(slot, node); }
int32_t orc_get_state(orc_forest* f, uint32_t slot, uint32_t node, spb_state* out) { return f->impl->get_state(slot, node, out); }
int32_t orc_arena_len(orc_forest* f, uint32_t slot, uint32_t* out) { return f->impl->arena_len(slot, out); }
int32_t orc_node_stats(orc_forest* f, uint32_t slot, uint32_t node, uint32_t* n, float* w, float* p, uint32_t* fc, uint32_t* nc) {
  return f->impl->node_stats(slot, node, n, w, p, fc, nc);
}
int32_t orc_get_counters(orc_forest* f, spb_counters* out) { f->impl->counters(out); return SPB_OK; }
int32_t orc_set_leaves_per_tree(orc_forest* f, uint32_t k) { if (k < 1 || k > 16) return SPB_ERR_ARG; f->impl->set_leaves_per_tree(k); return SPB_OK; }

int32_t orc_next_state(int32_t game, const spb_state* s, uint8_t action, spb_state* out) {
  if (game == SPB_GAME_CONNECT4) {
    C4State n;
    if (!C4State::from_abi(*s).get_next_state(action, &n)) return SPB_ERR_ILLEGAL;
    n.to_abi(out);
    return SPB_OK;
  }
  TttState n;
  if (!TttState::from_abi(*s).get_next_state(action, &n)) return SPB_ERR_ILLEGAL;
  n.to_abi(out);
  return SPB_OK;
}
uint32_t orc_valid_actions(int32_t game, const spb_state* s) {
  int acts[SPB_MAX_ACTIONS];
  int n = game == SPB_GAME_CONNECT4 ? C4State::from_abi(*s).get_valid_actions(acts) : TttState::from_abi(*s).get_valid_actions(acts);
  uint32_t m = 0;
  for (int i = 0; i < n; ++i) m |= 1u << acts[i];
  return m;
}
void orc_encode(int32_t game, const spb_state* s, float* out) {
  if (game == SPB_GAME_CONNECT4) C4State::from_abi(*s).get_encoding(out);
  else TttState::from_abi(*s).get_encoding(out);
}
void orc_mask_invalid_actions(int32_t game, const spb_state* s, const float* probs, float* out) {
  if (game == SPB_GAME_CONNECT4) C4State::from_abi(*s).mask_invalid_actions(probs, out);
  else TttState::from_abi(*s).mask_invalid_actions(probs, out);
}
void orc_det_eval(int32_t game, const spb_state* s, float* probs, float* value) {
  if (game == SPB_GAME_CONNECT4) det_eval(C4State::from_abi(*s), probs, value);
  else det_eval(TttState::from_abi(*s), probs, value);
}
uint64_t orc_det_hash(int32_t game, const spb_state* s) {
  return game == SPB_GAME_CONNECT4 ? det_hash(C4State::from_abi(*s)) : det_hash(TttState::from_abi(*s));
}
int32_t orc_greedy_game(int32_t game, const spb_state* root, float c, uint32_t num_searches, int32_t evaluator,
                        orc_eval_fn fn, void* user, uint8_t* actions, uint32_t* arena_sizes, uint32_t* last_counts,
                        uint32_t* n_last, uint8_t* final_status) {
  if (game == SPB_GAME_CONNECT4)
    return greedy_game<C4State>(root, c, num_searches, evaluator, fn, user, actions, arena_sizes, last_counts, n_last, final_status);
  return greedy_game<TttState>(root, c, num_searches, evaluator, fn, user, actions, arena_sizes, last_counts, n_last, final_status);
}
uint64_t orc_baseline_run(int32_t game, const spb_state* roots, uint32_t threads, uint32_t gpt, float c,
                          uint32_t num_searches, int32_t evaluator, orc_eval_fn fn, void* user, double* seconds) {
  if (game == SPB_GAME_CONNECT4) return baseline_run<C4State>(roots, threads, gpt, c, num_searches, evaluator, fn, user, seconds);
  return baseline_run<TttState>(roots, threads, gpt, c, num_searches, evaluator, fn, user, seconds);
}

}  // extern "C"
