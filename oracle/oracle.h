/*
 * oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (self-play-ai_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference (a Rust crate with un-vendored
 * dependencies) cannot be compiled in this image (no rustc/cargo) and ships no tests,
 * fixtures or golden vectors.  This restatement follows the reference line by line
 * (citations in oracle.cc) and is cross-checked against (a) the known-answer vectors of
 * SURVEY.md §8(c), produced by an independent numpy-f32 restatement, and (b) a second
 * independent pure-Python restatement in tests/pyref.py.
 */
#ifndef SPB_ORACLE_H
#define SPB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/selfplay_b200.h" /* spb_state, SPB_GAME_*, SPB_EVAL_* */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_forest orc_forest;

/*
 * Evaluator callback = the tensor part of Model::predict (model/mod.rs:60-67,95):
 * given n encodings [n][3][R][C] f32 it fills probs[n][A] (softmax output, NOT masked)
 * and values[n].  The oracle applies mask_invalid_actions itself.
 */
typedef void (*orc_eval_fn)(void* user, const float* encodings, uint32_t n, float* probs, float* values);

orc_forest* orc_create(int32_t game, uint32_t num_trees, float c);
void orc_destroy(orc_forest* f);
const char* orc_last_error(const orc_forest* f);

/* Tree::with_root_state for tree i (roots NULL = State::default()). */
int32_t orc_reset(orc_forest* f, const uint32_t* slots, uint32_t n, const spb_state* roots);
/* Mcts::search over all trees; evaluator = SPB_EVAL_DET / SPB_EVAL_UNIFORM built in, SPB_EVAL_NET via fn. */
int32_t orc_search(orc_forest* f, uint32_t num_searches, int32_t evaluator, orc_eval_fn fn, void* user);
int32_t orc_root_children(orc_forest* f, uint32_t slot, uint8_t* actions, uint32_t* visit_counts,
                          uint32_t* child_ids, uint32_t* n_children);
int32_t orc_root_policy(orc_forest* f, uint32_t slot, float* policy);
int32_t orc_use_subtree(orc_forest* f, uint32_t slot, uint32_t node_id);
int32_t orc_get_state(orc_forest* f, uint32_t slot, uint32_t node_id, spb_state* out);
int32_t orc_arena_len(orc_forest* f, uint32_t slot, uint32_t* out);
int32_t orc_node_stats(orc_forest* f, uint32_t slot, uint32_t node_id, uint32_t* visit_count,
                       float* value_sum, float* prior, uint32_t* first_child, uint32_t* n_children);
int32_t orc_get_counters(orc_forest* f, spb_counters* out);
/* EXTENSION (not in the reference): K in-flight leaves per tree per step with virtual loss; 1 = reference algorithm. */
int32_t orc_set_leaves_per_tree(orc_forest* f, uint32_t k);

/* State trait, one call per state (array-board restatement). */
int32_t orc_next_state(int32_t game, const spb_state* s, uint8_t action, spb_state* out);
uint32_t orc_valid_actions(int32_t game, const spb_state* s);
void orc_encode(int32_t game, const spb_state* s, float* out);
/* mask_invalid_actions: probs[A] -> out[A] (connect_four.rs:261-279, tictactoe.rs:218-236). */
void orc_mask_invalid_actions(int32_t game, const spb_state* s, const float* probs, float* out);
/* DetEval raw outputs. */
void orc_det_eval(int32_t game, const spb_state* s, float* probs, float* value);
uint64_t orc_det_hash(int32_t game, const spb_state* s);

/*
 * Greedy self-play of one game with subtree reuse (main.rs:106-114 move rule):
 * returns number of plies; actions[] (cap 64), final status; root_counts of the last searched
 * position scattered by CHILD order; arena sizes per move.
 */
int32_t orc_greedy_game(int32_t game, const spb_state* root, float c, uint32_t num_searches,
                        int32_t evaluator, orc_eval_fn fn, void* user, uint8_t* actions,
                        uint32_t* arena_sizes, uint32_t* last_counts, uint32_t* n_last, uint8_t* final_status);

/*
 * CPU baseline: `threads` workers, each an independent forest of `games_per_thread` trees rooted at
 * roots[t*games_per_thread + i], one search of num_searches (the shape of main.rs:169's SelfPlayWorkers).
 * Returns total simulations; *seconds receives wall time.  fn may be called concurrently from workers.
 */
uint64_t orc_baseline_run(int32_t game, const spb_state* roots, uint32_t threads, uint32_t games_per_thread,
                          float c, uint32_t num_searches, int32_t evaluator, orc_eval_fn fn, void* user,
                          double* seconds);

#ifdef __cplusplus
}
#endif
#endif
