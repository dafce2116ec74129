// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/algorithms/external_sampling_mccfr.{h,cc}: ExternalSamplingMCCFRSolver with
// AverageType::kSimple — RunIteration :71-80, UpdateRegrets :124-186 (sample chance and opponent nodes, walk every
// action at the traverser's nodes, regret += child value - node value, simple averaging at the next player's nodes),
// CFRInfoStateValues(legal, kInitialTableValues = 1e-6) external_sampling_mccfr.h:59, ApplyRegretMatching
// cfr.cc:596-615, SampleActionIndex(0, z) cfr.cc:617-628, SampleAction(outcomes, z) spiel.cc:372-409.
// Recursive, string-keyed, one State clone per edge, like the reference.
//
// Two sources for the uniform variates z:
//   rng_mode 0  std::mt19937(seed) + std::uniform_real_distribution<double>(0, 1), consumed in DFS order: the
//               reference's own stream (both are libstdc++ classes), so with traversals_per_update = 1 the tables
//               must equal the unmodified reference's bit for bit (tests/test_mccfr_oracle.py).
//   rng_mode 1  the device solver's injected stream: z = U53(Philox4x32-10(seed; path hash h, phase, traversal k)),
//               h(root) = 0, h(child) = h(parent) * 0x9E3779B97F4A7C15 + child index + 1, phase = iteration * players +
//               traverser, U53 = top 53 bits of (r1:r0) * 2^-53.  Keyed by position, not by consumption order.
// traversals_per_update = K: the K traversals of one (iteration, traverser) phase all read the tables as they were at
// the start of the phase; afterwards every table entry receives the sum of its K deltas, formed in the fixed order the
// device kernel k_mccfr_apply uses: 64 partial sums partial[q] = delta[q] + delta[q+64] + ... (sequential, from 0.0),
// the tree partial[q] += partial[q+s] for s = 32, 16, 8, 4, 2, 1, then table += partial[0].
// K = 1 is exactly the reference's algorithm (a traversal never revisits an information state it has updated).
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <random>
#include <string>
#include <vector>

#include "../oracle.h"
#include "philox.h"

namespace oracle {
namespace {

constexpr double kInitialTableValues = 0.000001;

struct McValues {
  std::vector<int64_t> legal;
  std::vector<double> regrets, cum_policy;
  int player = 0;
};

struct Delta { std::string key; bool average; std::vector<double> d; };

struct EsMccfr {
  const Game* game = nullptr;
  int n = 2, rng_mode = 0, K = 1, iteration = 0;
  bool full_average = false;       // AverageType::kFull (external_sampling_mccfr.h:53-54)
  uint64_t seed = 0;
  std::mt19937 mt;
  std::uniform_real_distribution<double> dist{0.0, 1.0};
  std::map<std::string, McValues> table;
  bool failed = false;

  double Z(uint64_t h, uint32_t phase, uint32_t k) {
    if (rng_mode == 0) return dist(mt);
    uint32_t r[4];
    Philox4(seed, h, phase, k, r);
    uint64_t bits = (((uint64_t)r[1] << 32) | r[0]) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
  }
  static uint64_t Child(uint64_t h, int idx) { return h * 0x9E3779B97F4A7C15ull + (uint64_t)(idx + 1); }

  McValues& Lookup(const State& s, int cur, const std::vector<int64_t>& la) {
    std::string key = s.InformationStateString(cur);
    auto it = table.find(key);
    if (it == table.end()) {
      McValues v;
      v.legal = la; v.player = cur;
      v.regrets.assign(la.size(), kInitialTableValues);
      v.cum_policy.assign(la.size(), kInitialTableValues);
      it = table.emplace(key, v).first;
    }
    return it->second;
  }

  double Update(const State& s, int player, uint64_t h, uint32_t phase, uint32_t k, std::vector<Delta>* out) {
    if (s.IsTerminal()) return s.Returns()[player];
    if (s.IsChanceNode()) {
      auto outcomes = s.ChanceOutcomes();
      double z = Z(h, phase, k);
      int chosen = -1;
      if (outcomes.size() == 1) chosen = 0;
      else {
        double sum = 0;
        for (size_t i = 0; i < outcomes.size(); ++i) {
          double prob = outcomes[i].second;
          if (sum <= z && z < sum + prob) { chosen = (int)i; break; }
          sum += prob;
        }
      }
      if (chosen < 0) { failed = true; chosen = (int)outcomes.size() - 1; }
      auto c = s.Clone();
      c->ApplyAction(outcomes[chosen].first);
      return Update(*c, player, Child(h, chosen), phase, k, out);
    }
    int cur = s.CurrentPlayer();
    auto la = s.LegalActions();
    const McValues& v = Lookup(s, cur, la);
    const size_t A = la.size();
    std::vector<double> policy(A);                               // ApplyRegretMatching on a copy
    {
      double sum_pos = 0.0;
      for (size_t a = 0; a < A; ++a) if (v.regrets[a] > 0) sum_pos += v.regrets[a];
      for (size_t a = 0; a < A; ++a)
        policy[a] = sum_pos > 0 ? (v.regrets[a] > 0 ? v.regrets[a] / sum_pos : 0) : 1.0 / A;
    }
    std::string key = s.InformationStateString(cur);
    double value = 0;
    std::vector<double> child_values(A, 0);
    if (cur != player) {
      double z = Z(h, phase, k);
      int aidx = -1;
      double sum = 0;
      for (size_t a = 0; a < A; ++a) {
        double prob = 0.0 * 1.0 / A + (1.0 - 0.0) * policy[a];
        if (z >= sum && z < sum + prob) { aidx = (int)a; break; }
        sum += prob;
      }
      if (aidx < 0) { failed = true; aidx = (int)A - 1; }
      auto c = s.Clone();
      c->ApplyAction(la[aidx]);
      value = Update(*c, player, Child(h, aidx), phase, k, out);
    } else {
      for (size_t a = 0; a < A; ++a) {
        auto c = s.Clone();
        c->ApplyAction(la[a]);
        child_values[a] = Update(*c, player, Child(h, (int)a), phase, k, out);
        value += policy[a] * child_values[a];
      }
    }
    if (cur == player) {
      Delta d{key, false, std::vector<double>(A)};
      for (size_t a = 0; a < A; ++a) d.d[a] = child_values[a] - value;
      out->push_back(d);
    }
    if (!full_average && cur == (player + 1) % n) out->push_back(Delta{key, true, policy});
    return value;
  }

  // FullUpdateAverage (external_sampling_mccfr.cc:188-230): one pass over the whole tree per iteration; every decision
  // node adds reach[current player] * regret-matching policy to its information state's cumulative policy (post-order)
  void FullUpdateAverage(const State& s, const std::vector<double>& reach) {
    if (s.IsTerminal()) return;
    if (s.IsChanceNode()) {
      for (auto a : s.LegalActions()) { auto c = s.Clone(); c->ApplyAction(a); FullUpdateAverage(*c, reach); }
      return;
    }
    double sum = 0.0;
    for (double r : reach) sum += r;
    if (sum == 0.0) return;
    int cur = s.CurrentPlayer();
    auto la = s.LegalActions();
    const size_t A = la.size();
    std::string key = s.InformationStateString(cur);
    std::vector<double> policy(A);
    {
      const McValues& v = Lookup(s, cur, la);
      double sum_pos = 0.0;
      for (size_t a = 0; a < A; ++a) if (v.regrets[a] > 0) sum_pos += v.regrets[a];
      for (size_t a = 0; a < A; ++a) policy[a] = sum_pos > 0 ? (v.regrets[a] > 0 ? v.regrets[a] / sum_pos : 0) : 1.0 / A;
    }
    for (size_t a = 0; a < A; ++a) {
      std::vector<double> nr = reach;
      nr[cur] *= policy[a];
      auto c = s.Clone();
      c->ApplyAction(la[a]);
      FullUpdateAverage(*c, nr);
    }
    McValues& v = table[key];
    for (size_t a = 0; a < A; ++a) v.cum_policy[a] += reach[cur] * policy[a];
  }

  void RunIteration() {
    for (int p = 0; p < n; ++p) {
      uint32_t phase = (uint32_t)(iteration * n + p);
      std::vector<std::vector<Delta>> deltas(K);
      for (int k = 0; k < K; ++k) {
        auto root = game->NewInitialState();
        Update(*root, p, 0, phase, (uint32_t)k, &deltas[k]);
      }
      // (key, average?) -> 64 partial sums per action
      std::map<std::pair<std::string, bool>, std::vector<std::vector<double>>> partial;
      for (int k = 0; k < K; ++k)
        for (const Delta& d : deltas[k]) {
          auto& ps = partial[{d.key, d.average}];
          if (ps.empty()) ps.assign(64, std::vector<double>(d.d.size(), 0.0));
          for (size_t a = 0; a < d.d.size(); ++a) if (d.d[a] != 0.0) ps[k % 64][a] += d.d[a];
        }
      for (auto& kv : partial) {
        auto& ps = kv.second;
        for (int s = 32; s >= 1; s >>= 1)
          for (int q = 0; q < s; ++q)
            for (size_t a = 0; a < ps[q].size(); ++a) ps[q][a] += ps[q + s][a];
        McValues& v = table[kv.first.first];
        for (size_t a = 0; a < ps[0].size(); ++a) (kv.first.second ? v.cum_policy : v.regrets)[a] += ps[0][a];
      }
    }
    if (full_average) {
      auto root = game->NewInitialState();
      FullUpdateAverage(*root, std::vector<double>(n, 1.0));
    }
    ++iteration;
  }
};

}  // namespace
}  // namespace oracle

extern "C" {

void orc_mccfr_set_full_average(void* m, int on) { ((oracle::EsMccfr*)m)->full_average = on != 0; }
void* orc_mccfr_new(void* game, uint64_t seed, int rng_mode, int traversals_per_update) {
  auto* m = new oracle::EsMccfr;
  m->game = (oracle::Game*)game;
  m->n = m->game->info.num_players;
  m->seed = seed; m->rng_mode = rng_mode; m->K = traversals_per_update;
  m->mt.seed((uint32_t)seed);
  return m;
}
void orc_mccfr_free(void* m) { delete (oracle::EsMccfr*)m; }
int orc_mccfr_iterate(void* m, int iters) {
  auto* s = (oracle::EsMccfr*)m;
  for (int i = 0; i < iters; ++i) s->RunIteration();
  return s->failed ? 1 : 0;
}
int orc_mccfr_num_infosets(void* m) { return (int)((oracle::EsMccfr*)m)->table.size(); }
// k-th entry in key order: key string, legal actions, cumulative regrets / policy; returns #actions.
int orc_mccfr_get(void* m, int k, char* key, int key_cap, int64_t* legal, double* regrets, double* cum, int cap, int* player) {
  auto& table = ((oracle::EsMccfr*)m)->table;
  auto it = table.begin();
  std::advance(it, k);
  strncpy(key, it->first.c_str(), key_cap - 1);
  key[key_cap - 1] = 0;
  int n = (int)it->second.legal.size();
  for (int i = 0; i < n && i < cap; ++i) {
    legal[i] = it->second.legal[i];
    regrets[i] = it->second.regrets[i];
    cum[i] = it->second.cum_policy[i];
  }
  if (player) *player = it->second.player;
  return n;
}

}  // extern "C"
