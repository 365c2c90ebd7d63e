// C++ host-side test of the reference-shaped API (self-play-ai_b200/host/selfplay_b200.hpp): plays the greedy game of
// main.rs:106-114 (arg-max visit count, last max wins, use_subtree) under DetEval and prints it for the pytest harness.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../self-play-ai_b200/host/selfplay_b200.hpp"

// chess through the C++ mirror: perft of the start position and a few greedy DetEval plies with use_subtree
static int chess_main(uint32_t sims) {
  try {
    spb::Args args;
    args.num_searches = sims;
    args.num_parallel_self_play_games = 1;
    spb::ChessMcts mcts(args, 0, SPB_EVAL_DET);
    spb::ChessState root = spb::ChessState::start();
    printf("perft3 %llu moves %zu\n", (unsigned long long)mcts.perft(root, 3), mcts.get_valid_actions(root).size());
    spb::ChessTree tree = mcts.with_root_state(0, root);
    std::vector<spb::ChessTree*> trees{&tree};
    printf("moves");
    for (int ply = 0; ply < 4; ++ply) {
      auto res = mcts.search(trees);
      const auto& pairs = res[0].child_id_to_probs;
      size_t best = 0;
      for (size_t i = 1; i < pairs.size(); ++i)
        if (!(pairs[best].second > pairs[i].second)) best = i;
      printf(" %u", (unsigned)res[0].moves[best]);
      fprintf(stderr, "ply %d arena %zu\n", ply, mcts.arena_len(tree));
      mcts.use_subtree(tree, pairs[best].first);
    }
    printf("\n");
    try { mcts.use_subtree(tree, 0); printf("no-throw\n"); } catch (const spb::Error& e) { printf("error %d\n", e.code); }
  } catch (const spb::Error& e) {
    printf("FAILED %s\n", e.what());
    return 1;
  }
  return 0;
}

int main(int argc, char** argv) {
  const uint32_t sims = argc > 1 ? (uint32_t)atoi(argv[1]) : 800;
  if (argc > 2 && std::string(argv[2]) == "chess") return chess_main(sims);
  try {
    spb::Args args;
    args.num_searches = sims;
    args.num_parallel_self_play_games = 1;
    spb::Mcts mcts(args, SPB_GAME_CONNECT4, 0, SPB_EVAL_DET);
    spb::Tree tree = mcts.make_tree(0);
    std::vector<spb::Tree*> trees{&tree};
    printf("actions");
    for (int ply = 0; ply < 42; ++ply) {
      auto res = mcts.search(trees);
      const auto& pairs = res[0].second;
      size_t best = 0;
      for (size_t i = 1; i < pairs.size(); ++i)
        if (!(pairs[best].second > pairs[i].second)) best = i;          // max_by(total_cmp): last max wins
      printf(" %d", (int)mcts.action_taken(tree, best));
      size_t arena = mcts.arena_len(tree);
      spb::State s = mcts.use_subtree(tree, pairs[best].first);
      fprintf(stderr, "ply %d arena %zu\n", ply, arena);
      if (s.get_status() != spb::Status::Ongoing) { printf("\nstatus %d\n", (int)s.raw.status); break; }
    }
    // error behaviour: out-of-range node id is an error, not a crash
    try { mcts.node_state(tree, 1u << 23); printf("no-throw\n"); } catch (const spb::Error& e) { printf("error %d\n", e.code); }
  } catch (const spb::Error& e) {
    printf("FAILED %s\n", e.what());
    return 1;
  }
  return 0;
}
