// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/algorithms/mcts.{h,cc}: MCTSBot::MCTSearch with
// RandomRolloutEvaluator, UCT child selection and the MCTS-Solver back-propagation, for deterministic
// perfect-information games.  Structure follows the reference (ApplyTreePolicy :273-351, MCTSearch :353-467,
// UCTValue :90-101, BestChild/CompareFinal :114-143, RandomRolloutEvaluator::Evaluate :43-72).
// The reference draws from two std::mt19937 streams through std::shuffle and absl::Uniform, whose bit patterns
// are library specific (RNG-stream parity unpinned); here every random decision is an explicit
// function of (seed key, tree, simulation / expansion index, position) through Philox (philox.h), and the CUDA
// kernels consume the identical function, so the search trees can be compared bit for bit:
//   * expansion #e of a tree: Fisher-Yates over the ascending legal-action list, for i = n-1..1:
//       j = RngUniform(key, e, i, 1, i + 1); swap(list[i], list[j])          (replaces std::shuffle, mcts.cc:294)
//   * simulation #t, rollout #r, rollout ply p: draw q = 0, 1, ...: k = RngUniformShared(key, t, p + 4096 q, 2 + r, C)
//       (word (b & 3) of the Philox block of (t, b >> 2, 2 + r): four consecutive plies share one block)
//       over the C rollout candidates (State::RolloutCandidates, = the legal actions except for go); the first
//       legal candidate is played — a uniform draw over the legal actions    (replaces absl::Uniform, mcts.cc:54)
//   key = seed + tree_index * 0x9E3779B97F4A7C15.
#include <cmath>
#include <algorithm>
#include <limits>
#include <random>

#include "../oracle.h"
#include "philox.h"

namespace oracle {
namespace {

struct Node {                      // SearchNode, mcts.h:114-146
  int64_t action = kInvalidAction;
  int player = 0;                  // the player who chose `action`
  int explore_count = 0;
  double total_reward = 0;
  std::vector<double> outcome;
  std::vector<Node> children;
};

double UctValue(const Node& n, int parent_explore_count, double uct_c) {   // mcts.cc:90-101
  if (!n.outcome.empty()) return n.outcome[n.player];
  if (n.explore_count == 0) return std::numeric_limits<double>::infinity();
  return n.total_reward / n.explore_count + uct_c * std::sqrt(std::log(parent_explore_count) / n.explore_count);
}

double PuctValue(const Node& n, int parent_explore_count, double uct_c, double prior) {   // mcts.cc:103-112
  if (!n.outcome.empty()) return n.outcome[n.player];
  return ((n.explore_count != 0 ? n.total_reward / n.explore_count : 0) +
          uct_c * prior * std::sqrt(parent_explore_count) / (n.explore_count + 1));
}

bool CompareFinal(const Node& a, const Node& b) {                          // mcts.cc:114-125
  double out = (a.player >= 0 && a.player < (int)a.outcome.size()) ? a.outcome[a.player] : 0;
  double out_b = (b.player >= 0 && b.player < (int)b.outcome.size()) ? b.outcome[b.player] : 0;
  if (out != out_b) return out < out_b;
  if (a.explore_count != b.explore_count) return a.explore_count < b.explore_count;
  return a.total_reward < b.total_reward;
}

// Uniform random legal action by rejection from the candidate list; draw(b, n) supplies the random integers.
template <typename Draw>
int64_t SampleRolloutAction(const State& s, Draw draw, uint32_t ply) {
  auto cand = s.RolloutCandidates();
  auto legal = s.LegalActions();
  for (uint32_t retry = 0;; ++retry) {
    int64_t a = cand[draw(ply + 4096u * retry, (uint32_t)cand.size())];
    for (auto l : legal) if (l == a) return a;
  }
}

// absl::Uniform(rng, 0u, n) as the reference calls it at mcts.cc:54 (common type of unsigned and size_t = a 64-bit
// unsigned): abseil's published uniform_int_distribution algorithm — a 64-bit word assembled from two mt19937 draws
// (first draw in the high half), a mask when n is a power of two, else Lemire's multiply-shift with rejection of the
// low products below (2^64 - n) mod n.  oracle/absl_shim implements the same published algorithm for the reference
// build; seeded parity with stock abseil binaries stays unpinned.
uint64_t AbslUniformBelow(std::mt19937& g, uint64_t n) {
  auto bits64 = [&]() { uint64_t hi = g(); return (hi << 32) + (uint64_t)g(); };
  uint64_t bits = bits64();
  const uint64_t r = n - 1;
  if ((r & n) == 0) return bits & r;
  unsigned __int128 product = (unsigned __int128)bits * n;
  if ((uint64_t)product < n) {
    const uint64_t threshold = (0 - n) % n;
    while ((uint64_t)product < threshold) { bits = bits64(); product = (unsigned __int128)bits * n; }
  }
  return (uint64_t)(product >> 64);
}

struct Search {
  // rng_mode 0: the position-keyed Philox stream shared with the device kernel (default).
  // rng_mode 1: the reference's own streams — MCTSBot::rng_ (std::mt19937(seed), std::shuffle of new children,
  //             mcts.cc:294) and RandomRolloutEvaluator::rng_ (std::mt19937(seed), absl::Uniform over LegalActions(),
  //             mcts.cc:54) — so that results can be compared with the unmodified reference bit for bit.
  int rng_mode = 0;
  std::mt19937 bot_rng, eval_rng;
  uint64_t key;
  double uct_c, max_utility;
  bool puct = false;               // ChildSelectionPolicy (mcts.h:148)
  int n_rollouts;
  bool solve;
  uint32_t expansions = 0;
  int nodes = 1;                   // MCTSBot::nodes_ (mcts.h:209)
  int max_nodes = 1;               // MCTSBot::max_nodes_ = (max_memory_mb << 20) / sizeof(SearchNode) + 1; <= 1: never collect
  int gc_limit = 5;                // MCTSBot::gc_limit_, starts at MIN_GC_LIMIT (mcts.cc:37, 355)
  int gc_runs = 0;

  // MCTSBot::GarbageCollect (mcts.cc:469-482): drop the children of every node explored fewer than gc_limit times
  void GarbageCollect(Node* node) {
    if (node->children.empty()) return;
    bool clear_children = node->explore_count < gc_limit;
    for (Node& child : node->children) GarbageCollect(&child);
    if (clear_children) {
      nodes -= (int)node->children.capacity();
      node->children.clear();
      node->children.shrink_to_fit();
    }
  }

  std::vector<double> Evaluate(const State& state, uint32_t sim) {        // mcts.cc:43-72
    std::vector<double> result;
    for (int r = 0; r < n_rollouts; ++r) {
      auto ws = state.Clone();
      uint32_t ply = 0;
      while (!ws->IsTerminal()) {
        if (rng_mode == 1) {
          auto actions = ws->LegalActions();
          ws->ApplyAction(actions[AbslUniformBelow(eval_rng, actions.size())]);
        } else {
          ws->ApplyAction(SampleRolloutAction(*ws, [&](uint32_t b, uint32_t n) { return RngUniformShared(key, sim, b, 2 + r, n); }, ply));
        }
        ++ply;
      }
      auto returns = ws->Returns();
      if (result.empty()) result.swap(returns);
      else for (size_t i = 0; i < result.size(); ++i) result[i] += returns[i];
    }
    for (auto& v : result) v /= n_rollouts;
    return result;
  }

  std::unique_ptr<State> TreePolicy(Node* root, const State& state, std::vector<Node*>* path) {   // mcts.cc:273-351
    path->push_back(root);
    auto ws = state.Clone();
    Node* cur = root;
    while (!ws->IsTerminal() && cur->explore_count > 0) {
      if (cur->children.empty()) {
        auto legal = ws->LegalActions();             // uniform prior over LegalActions (mcts.cc:74-87)
        uint32_t e = expansions++;
        if (rng_mode == 1) std::shuffle(legal.begin(), legal.end(), bot_rng);
        else for (int i = (int)legal.size() - 1; i >= 1; --i) std::swap(legal[i], legal[RngUniform(key, e, i, 1, i + 1)]);
        int player = ws->CurrentPlayer();
        cur->children.reserve(legal.size());
        for (auto a : legal) { Node c; c.action = a; c.player = player; cur->children.push_back(c); }
        nodes += (int)cur->children.capacity();
      }
      Node* chosen = nullptr;
      double max_value = -std::numeric_limits<double>::infinity();
      for (Node& child : cur->children) {
        // the prior is RandomRolloutEvaluator::Prior's 1.0 / legal_actions.size() (mcts.cc:74-87)
        double val = puct ? PuctValue(child, cur->explore_count, uct_c, 1.0 / cur->children.size())
                          : UctValue(child, cur->explore_count, uct_c);
        if (val > max_value) { max_value = val; chosen = &child; }
      }
      cur = chosen;
      ws->ApplyAction(chosen->action);
      path->push_back(cur);
    }
    return ws;
  }

  // returns the number of simulations actually run (early exit when the root is solved / has one child)
  int Run(Node* root, const State& state, int max_simulations) {          // mcts.cc:353-467
    std::vector<Node*> path;
    int i = 0;
    for (; i < max_simulations; ++i) {
      path.clear();
      auto ws = TreePolicy(root, state, &path);
      std::vector<double> returns;
      bool solved;
      if (ws->IsTerminal()) {
        returns = ws->Returns();
        path.back()->outcome = returns;
        solved = solve;
      } else {
        returns = Evaluate(*ws, (uint32_t)i);
        solved = false;
      }
      while (!path.empty()) {
        Node* node = path.back();
        node->total_reward += returns[node->player];
        node->explore_count += 1;
        path.pop_back();
        if (solved && !node->children.empty()) {
          int player = node->children[0].player;
          const Node* best = nullptr;
          bool all_solved = true;
          for (const Node& child : node->children) {
            if (child.outcome.empty()) all_solved = false;
            else if (best == nullptr || child.outcome[player] > best->outcome[player]) best = &child;
          }
          if (best != nullptr && (all_solved || best->outcome[player] == max_utility)) node->outcome = best->outcome;
          else solved = false;
        }
      }
      if (!root->outcome.empty() || root->children.size() == 1) { ++i; break; }
      if (max_nodes > 1 && nodes >= max_nodes) {                          // mcts.cc:441-463
        GarbageCollect(root);
        ++gc_runs;
        gc_limit *= (nodes > max_nodes / 2 ? 1.25 : 0.9);                 // int *= double, as the reference's int gc_limit_
        gc_limit = std::max(5, gc_limit);
      }
    }
    return i;
  }
};

}  // namespace
}  // namespace oracle

extern "C" {

// One MCTSearch from `state` for tree #tree_index.  Root children are reported in child (shuffled) order:
// child_actions/visits/rewards/outcome_p0 (NaN when unproven).  Returns the number of root children.
// max_nodes = MCTSBot::max_nodes_ (<= 1: no garbage collection); gc_runs_out (nullable) = collections performed.
int orc_mcts_search_gc(void* game, void* state, double uct_c, int max_simulations, int n_rollouts, int solve,
                       uint64_t seed, uint64_t tree_index, int64_t* child_actions, int* child_visits,
                       double* child_rewards, double* child_outcome_p0, int cap, int64_t* best_action,
                       int* root_visits, double* root_outcome_p0, long* nodes_out, int* sims_run,
                       int child_selection_policy, int rng_mode, int max_nodes, int* gc_runs_out) {
  using namespace oracle;
  Game* g = (Game*)game;
  State* s = (State*)state;
  Search srch;
  srch.key = seed + tree_index * 0x9E3779B97F4A7C15ull;
  srch.uct_c = uct_c;
  srch.max_utility = g->info.max_utility;
  srch.n_rollouts = n_rollouts;
  srch.solve = solve != 0;
  srch.puct = child_selection_policy == 1;
  srch.rng_mode = rng_mode;
  srch.max_nodes = max_nodes;
  if (rng_mode == 1) { srch.bot_rng.seed((uint32_t)seed); srch.eval_rng.seed((uint32_t)seed); }
  Node root;
  root.player = s->CurrentPlayer();
  int ran = srch.Run(&root, *s, max_simulations);
  int n = (int)root.children.size();
  for (int i = 0; i < n && i < cap; ++i) {
    child_actions[i] = root.children[i].action;
    child_visits[i] = root.children[i].explore_count;
    child_rewards[i] = root.children[i].total_reward;
    child_outcome_p0[i] = root.children[i].outcome.empty() ? std::nan("") : root.children[i].outcome[0];
  }
  if (best_action) {
    *best_action = kInvalidAction;
    if (n) {
      const Node* best = &root.children[0];
      for (int i = 1; i < n; ++i) if (CompareFinal(*best, root.children[i])) best = &root.children[i];   // std::max_element
      *best_action = best->action;
    }
  }
  if (root_visits) *root_visits = root.explore_count;
  if (root_outcome_p0) *root_outcome_p0 = root.outcome.empty() ? std::nan("") : root.outcome[0];
  if (nodes_out) *nodes_out = srch.nodes;
  if (sims_run) *sims_run = ran;
  if (gc_runs_out) *gc_runs_out = srch.gc_runs;
  return n;
}

int orc_mcts_search(void* game, void* state, double uct_c, int max_simulations, int n_rollouts, int solve,
                    uint64_t seed, uint64_t tree_index, int64_t* child_actions, int* child_visits,
                    double* child_rewards, double* child_outcome_p0, int cap, int64_t* best_action,
                    int* root_visits, double* root_outcome_p0, long* nodes_out, int* sims_run,
                    int child_selection_policy, int rng_mode) {
  return orc_mcts_search_gc(game, state, uct_c, max_simulations, n_rollouts, solve, seed, tree_index, child_actions, child_visits,
                            child_rewards, child_outcome_p0, cap, best_action, root_visits, root_outcome_p0, nodes_out, sims_run,
                            child_selection_policy, rng_mode, 1, nullptr);
}

}  // extern "C"
