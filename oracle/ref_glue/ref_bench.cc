// TEST INFRASTRUCTURE ONLY.  CPU timing harness over the UNMODIFIED reference (oracle/_ref/ref_bench), used by
// bench.py --impl reference / cpu_baseline.  Mirrors oracle/algorithms/bench_apply.cc but steps real
// open_spiel::State objects:  ref_bench apply <game> <n> <max_prefix> <seed> <threads> <reps>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "open_spiel/algorithms/cfr.h"
#include "open_spiel/algorithms/external_sampling_mccfr.h"
#include "open_spiel/algorithms/mcts.h"
#include "open_spiel/spiel.h"

using open_spiel::Action;
using open_spiel::State;

static int BenchApply(const std::string& game_str, long n, int max_prefix, unsigned long seed, int threads, int reps) {
  auto game = open_spiel::LoadGame(game_str);
  std::vector<std::vector<double>> rep_secs(threads, std::vector<double>(reps, 0.0));
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      long lo = n * t / threads, hi = n * (t + 1) / threads;
      std::mt19937_64 rng(seed + 977 * t);
      std::vector<std::unique_ptr<State>> base;
      std::vector<Action> act;
      for (long i = lo; i < hi; ++i) {
        for (;;) {
          auto s = game->NewInitialState();
          int k = (int)(rng() % (unsigned long)(max_prefix + 1));
          for (int j = 0; j < k && !s->IsTerminal(); ++j) {
            auto la = s->LegalActions();
            s->ApplyAction(la[rng() % la.size()]);
          }
          if (s->IsTerminal()) continue;
          auto la = s->LegalActions();
          act.push_back(la[rng() % la.size()]);
          base.push_back(std::move(s));
          break;
        }
      }
      for (int r = 0; r < reps; ++r) {
        std::vector<std::unique_ptr<State>> work;
        work.reserve(base.size());
        for (auto& s : base) work.push_back(s->Clone());
        auto t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < work.size(); ++i) work[i]->ApplyAction(act[i]);
        auto t1 = std::chrono::steady_clock::now();
        rep_secs[t][r] = std::chrono::duration<double>(t1 - t0).count();
      }
    });
  }
  for (auto& th : pool) th.join();
  double total = 0;
  std::string per = "[";
  for (int r = 0; r < reps; ++r) {
    double m = 0;
    for (int t = 0; t < threads; ++t) m = rep_secs[t][r] > m ? rep_secs[t][r] : m;
    total += m;
    char b[64];
    snprintf(b, sizeof b, "%s%.9g", r ? "," : "", m);
    per += b;
  }
  per += "]";
  printf("{\"steps_per_s\": %.6g, \"seconds\": %.9g, \"per_rep_seconds\": %s, \"threads\": %d}\n",
         total > 0 ? (double)n * reps / total : 0.0, total, per.c_str(), threads);
  return 0;
}

// ref_bench mcts <game> <sims> <seed> <threads>: one MCTSearch of <sims> simulations per thread from the initial state.
static int BenchMcts(const std::string& game_str, int sims, int seed, int threads) {
  auto game = open_spiel::LoadGame(game_str);
  std::vector<double> secs(threads, 0.0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      auto evaluator = std::make_shared<open_spiel::algorithms::RandomRolloutEvaluator>(1, seed + t);
      open_spiel::algorithms::MCTSBot bot(*game, evaluator, 2.0, sims, 1000, true, seed + t, false);
      auto state = game->NewInitialState();
      auto t0 = std::chrono::steady_clock::now();
      auto root = bot.MCTSearch(*state);
      auto t1 = std::chrono::steady_clock::now();
      secs[t] = std::chrono::duration<double>(t1 - t0).count();
    });
  }
  for (auto& th : pool) th.join();
  double mx = 0;
  for (double s : secs) mx = s > mx ? s : mx;
  printf("{\"sims_per_s\": %.6g, \"seconds\": %.9g, \"threads\": %d, \"sims_per_tree\": %d}\n",
         mx > 0 ? (double)sims * threads / mx : 0.0, mx, threads, sims);
  return 0;
}

// ref_bench cfr <game> <iters>
static int BenchCfr(const std::string& game_str, int iters) {
  auto game = open_spiel::LoadGame(game_str);
  open_spiel::algorithms::CFRSolver solver(*game);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; ++i) solver.EvaluateAndUpdatePolicy();
  auto t1 = std::chrono::steady_clock::now();
  double s = std::chrono::duration<double>(t1 - t0).count();
  printf("{\"iters_per_s\": %.6g, \"seconds\": %.9g, \"iters\": %d}\n", s > 0 ? iters / s : 0.0, s, iters);
  return 0;
}

// ref_bench rollout <game> <games> <seed> <threads>: uniform-random playouts to the end, the loop of
// examples/benchmark_game.cc:32-115 (LegalActions + ApplyAction per ply; no observation tensor), <games> per thread.
static int BenchRollout(const std::string& game_str, long games, int seed, int threads) {
  auto game = open_spiel::LoadGame(game_str);
  std::vector<double> secs(threads, 0.0);
  std::vector<long> plies(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      std::mt19937 rng(seed + t);
      auto t0 = std::chrono::steady_clock::now();
      long moves = 0;
      for (long g = 0; g < games; ++g) {
        auto s = game->NewInitialState();
        while (!s->IsTerminal()) {
          auto la = s->LegalActions();
          std::uniform_int_distribution<int> pick(0, (int)la.size() - 1);
          s->ApplyAction(la[pick(rng)]);
          ++moves;
        }
      }
      auto t1 = std::chrono::steady_clock::now();
      secs[t] = std::chrono::duration<double>(t1 - t0).count();
      plies[t] = moves;
    });
  }
  for (auto& th : pool) th.join();
  double mx = 0;
  long total_plies = 0;
  for (int t = 0; t < threads; ++t) { mx = secs[t] > mx ? secs[t] : mx; total_plies += plies[t]; }
  printf("{\"games_per_s\": %.6g, \"plies_per_s\": %.6g, \"seconds\": %.9g, \"threads\": %d, \"games\": %ld}\n",
         mx > 0 ? (double)games * threads / mx : 0.0, mx > 0 ? total_plies / mx : 0.0, mx, threads, games * threads);
  return 0;
}

// ref_bench mccfr <game> <iters> <seed>: ExternalSamplingMCCFRSolver::RunIteration, one thread (the algorithm is sequential).
static int BenchMccfr(const std::string& game_str, int iters, int seed) {
  auto game = open_spiel::LoadGame(game_str);
  open_spiel::algorithms::ExternalSamplingMCCFRSolver solver(*game, seed);
  for (int i = 0; i < 50; ++i) solver.RunIteration();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; ++i) solver.RunIteration();
  auto t1 = std::chrono::steady_clock::now();
  double s = std::chrono::duration<double>(t1 - t0).count();
  printf("{\"traversals_per_s\": %.6g, \"iters_per_s\": %.6g, \"seconds\": %.9g, \"iters\": %d}\n",
         s > 0 ? 2.0 * iters / s : 0.0, s > 0 ? iters / s : 0.0, s, iters);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 6 && !strcmp(argv[1], "rollout")) return BenchRollout(argv[2], atol(argv[3]), atoi(argv[4]), atoi(argv[5]));
  if (argc >= 5 && !strcmp(argv[1], "mccfr")) return BenchMccfr(argv[2], atoi(argv[3]), atoi(argv[4]));
  if (argc >= 8 && !strcmp(argv[1], "apply"))
    return BenchApply(argv[2], atol(argv[3]), atoi(argv[4]), strtoul(argv[5], nullptr, 10), atoi(argv[6]), atoi(argv[7]));
  if (argc >= 6 && !strcmp(argv[1], "mcts")) return BenchMcts(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  if (argc >= 4 && !strcmp(argv[1], "cfr")) return BenchCfr(argv[2], atoi(argv[3]));
  fprintf(stderr, "usage: ref_bench apply|mcts|cfr ...\n");
  return 2;
}
