// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  CPU timing harness for bench.py's cpu_baseline leg:
// the reference's way of stepping states (one heap State per position, ApplyAction through a virtual
// call — spiel.cc:441-451) replayed over a synthetic (state, action) stream, sharded over host threads.
#include <chrono>
#include <random>
#include <thread>

#include "../oracle.h"

using oracle::Game;
using oracle::State;

extern "C" {

// Builds n states by advancing each k_i ~ U{0..max_prefix} uniformly random legal plies from the initial
// state (non-terminal states only), picks one uniformly random legal action for each, then times
// `reps` passes of: clone all states (untimed), ApplyAction on every state (timed).  Returns total
// ApplyAction calls per second of timed region (max over threads of the per-thread time).
double orc_bench_apply(void* game, int64_t n, int max_prefix, uint64_t seed, int threads, int reps,
                       double* seconds_out, double* per_rep_seconds /* [reps], nullable */) {
  Game* g = (Game*)game;
  if (threads < 1) threads = 1;
  std::vector<double> secs(threads, 0.0);
  std::vector<std::vector<double>> rep_secs(threads, std::vector<double>(reps, 0.0));
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
      std::mt19937_64 rng(seed + 977 * t);
      std::vector<std::unique_ptr<State>> base;
      std::vector<int64_t> act;
      base.reserve(hi - lo);
      for (int64_t i = lo; i < hi; ++i) {
        for (;;) {
          auto s = g->NewInitialState();
          int k = (int)(rng() % (uint64_t)(max_prefix + 1));
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
      double total = 0;
      for (int r = 0; r < reps; ++r) {
        std::vector<std::unique_ptr<State>> work;
        work.reserve(base.size());
        for (auto& s : base) work.push_back(s->Clone());
        auto t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < work.size(); ++i) work[i]->ApplyAction(act[i]);
        auto t1 = std::chrono::steady_clock::now();
        rep_secs[t][r] = std::chrono::duration<double>(t1 - t0).count();
        total += rep_secs[t][r];
      }
      secs[t] = total;
    });
  }
  for (auto& th : pool) th.join();
  double mx = 0;
  for (double s : secs) mx = s > mx ? s : mx;
  if (seconds_out) *seconds_out = mx;
  if (per_rep_seconds)
    for (int r = 0; r < reps; ++r) {
      double m = 0;
      for (int t = 0; t < threads; ++t) m = rep_secs[t][r] > m ? rep_secs[t][r] : m;
      per_rep_seconds[r] = m;
    }
  return mx > 0 ? (double)n * reps / mx : 0.0;
}

}  // extern "C"
