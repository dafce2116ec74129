// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/algorithms/outcome_sampling_mccfr.{h,cc}: OutcomeSamplingMCCFRSolver with the
// default uniform policy and no baseline — RunIteration :60-67 (one sampled episode per player), SampleEpisode :150-247
// (epsilon-on-policy sampling at the update player's nodes :139-147, on-policy at the others, importance-weighted
// value estimate, regret and average-policy updates at the update player's nodes), CFRInfoStateValues(legal,
// kInitialTableValues = 1e-6), ApplyRegretMatching cfr.cc:596-615.  Recursive and string-keyed like the reference.
//
// Two sources of randomness:
//   rng_mode 0  the reference's own stream as built here: std::mt19937(seed) consumed by absl::uniform_real_distribution
//               (chance nodes, SampleAction :157) and absl::discrete_distribution (actions, :185-187) — the published
//               algorithms as oracle/absl_shim states them (64 bits from two engine draws, first draw high; u = (bits >> 11)
//               * 2^-53; discrete: u * total, first index with u < running sum).  With trajectories_per_update = 1 the
//               tables equal the shim-built unmodified reference bit for bit (tests/test_os_mccfr_oracle.py); seeded
//               parity with binaries built against stock abseil stays unpinned.
//   rng_mode 1  the device solver's injected stream: z = U53(Philox4x32-10(seed; path hash h, phase, trajectory k)) with the
//               same hash / phase definitions as oracle/algorithms/mccfr.cc; both samplers use z the way the shim uses u.
// trajectories_per_update = K: the K episodes of one (iteration, player) phase read the tables as they were at the start of
// the phase; every entry then receives the sum of its K deltas in the fixed order of the device kernel (64 partial sums
// partial[q] = delta[q] + delta[q+64] + ..., the tree partial[q] += partial[q+s] for s = 32..1, table += partial[0]).
// K = 1 is exactly the reference's algorithm.
#include <cstdint>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "../oracle.h"
#include "philox.h"

namespace oracle {
namespace {

constexpr double kInitialTableValues = 0.000001;

struct OsValues {
  std::vector<int64_t> legal;
  std::vector<double> regrets, cum_policy;
  int player = 0;
};
struct OsDelta { std::string key; std::vector<double> regret, policy; };

struct OsMccfr {
  const Game* game = nullptr;
  int n = 2, rng_mode = 0, K = 1, iteration = 0;
  double epsilon = 0.6;
  uint64_t seed = 0;
  std::mt19937 mt;
  std::map<std::string, OsValues> table;
  bool failed = false;

  double U53() {                     // shim FastBits<uint64_t> + uniform_real_distribution<double>(0, 1)
    uint64_t hi = mt(), lo = mt();
    uint64_t bits = (hi << 32) + lo;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
  }
  // uniform_real_distribution<double>(lo, hi): lo + u (hi - lo), redrawn while the result rounds up to hi
  double Real(double lo, double hi, uint64_t h, uint32_t phase, uint32_t k, uint32_t* redraw) {
    for (;;) {
      double u;
      if (rng_mode == 0) u = U53();
      else {
        uint32_t r[4];
        Philox4(seed, h + 0x632BE59BD9B4E019ull * (uint64_t)(*redraw), phase, k, r);
        ++*redraw;
        u = (double)((((uint64_t)r[1] << 32) | r[0]) >> 11) * (1.0 / 9007199254740992.0);
      }
      double r = lo + u * (hi - lo);
      if (r < hi || lo == hi) return r;
    }
  }
  static uint64_t Child(uint64_t h, int idx) { return h * 0x9E3779B97F4A7C15ull + (uint64_t)(idx + 1); }

  OsValues& Lookup(const State& s, int cur, const std::vector<int64_t>& la) {
    std::string key = s.InformationStateString(cur);
    auto it = table.find(key);
    if (it == table.end()) {
      OsValues v;
      v.legal = la; v.player = cur;
      v.regrets.assign(la.size(), kInitialTableValues);
      v.cum_policy.assign(la.size(), kInitialTableValues);
      it = table.emplace(key, v).first;
    }
    return it->second;
  }

  double Episode(State* s, int update_player, uint64_t h, uint32_t phase, uint32_t k, double my_reach, double opp_reach,
                 double sample_reach, std::vector<OsDelta>* out) {
    if (s->IsTerminal()) return s->Returns()[update_player];
    uint32_t redraw = 0;
    if (s->IsChanceNode()) {
      auto outcomes = s->ChanceOutcomes();
      double z = Real(0.0, 1.0, h, phase, k, &redraw);
      int chosen = -1;                                  // SampleAction(outcomes, z), spiel.cc:372-409
      double sum = 0;
      for (size_t i = 0; i < outcomes.size(); ++i) {
        double prob = outcomes[i].second;
        if (sum <= z && z < sum + prob) { chosen = (int)i; break; }
        sum += prob;
      }
      if (chosen < 0) { failed = true; chosen = (int)outcomes.size() - 1; }
      double prob = outcomes[chosen].second;
      s->ApplyAction(outcomes[chosen].first);
      return Episode(s, update_player, Child(h, chosen), phase, k, my_reach, prob * opp_reach, prob * sample_reach, out);
    }
    int player = s->CurrentPlayer();
    auto la = s->LegalActions();
    const size_t A = la.size();
    std::string key = s->InformationStateString(player);
    std::vector<double> policy(A);                      // info_state_copy.ApplyRegretMatching()
    {
      const OsValues& v = Lookup(*s, player, la);
      double sum_pos = 0.0;
      for (size_t a = 0; a < A; ++a) if (v.regrets[a] > 0) sum_pos += v.regrets[a];
      for (size_t a = 0; a < A; ++a) policy[a] = sum_pos > 0 ? (v.regrets[a] > 0 ? v.regrets[a] / sum_pos : 0) : 1.0 / A;
    }
    std::vector<double> sample_policy = policy;
    if (player == update_player)
      for (size_t a = 0; a < A; ++a) sample_policy[a] = epsilon * 1.0 / A + (1 - epsilon) * policy[a];     // SamplePolicy :139-147
    double total = 0.0;                                 // absl::discrete_distribution(sample_policy)
    for (double w : sample_policy) total += w;
    double u = Real(0.0, total, h, phase, k, &redraw), acc = 0;
    int sampled = (int)A - 1;
    for (size_t a = 0; a < A; ++a) { acc += sample_policy[a]; if (u < acc) { sampled = (int)a; break; } }
    s->ApplyAction(la[sampled]);
    double child_value = Episode(s, update_player, Child(h, sampled), phase, k,
                                 player == update_player ? my_reach * policy[sampled] : my_reach,
                                 player == update_player ? opp_reach : opp_reach * policy[sampled],
                                 sample_reach * sample_policy[sampled], out);
    std::vector<double> child_values(A, 0.0);           // BaselineCorrectedChildValue with baseline 0 (:128-137)
    for (size_t a = 0; a < A; ++a) child_values[a] = (int)a == sampled ? 0.0 + (child_value - 0.0) / sample_policy[a] : 0.0;
    double value_estimate = 0;
    for (size_t a = 0; a < A; ++a) value_estimate += policy[a] * child_values[a];
    if (player == update_player) {
      double cf_value = value_estimate * opp_reach / sample_reach;
      OsDelta d{key, std::vector<double>(A), std::vector<double>(A)};
      for (size_t a = 0; a < A; ++a) {
        double cf_action_value = child_values[a] * opp_reach / sample_reach;
        d.regret[a] = cf_action_value - cf_value;
        d.policy[a] = my_reach * policy[a] / sample_reach;
      }
      out->push_back(d);
    }
    return value_estimate;
  }

  void RunIteration() {
    for (int p = 0; p < n; ++p) {
      uint32_t phase = (uint32_t)(iteration * n + p);
      std::vector<std::vector<OsDelta>> deltas(K);
      for (int k = 0; k < K; ++k) {
        auto root = game->NewInitialState();
        Episode(root.get(), p, 0, phase, (uint32_t)k, 1.0, 1.0, 1.0, &deltas[k]);
      }
      std::map<std::pair<std::string, bool>, std::vector<std::vector<double>>> partial;   // (key, policy?) -> 64 lanes
      for (int k = 0; k < K; ++k)
        for (const OsDelta& d : deltas[k])
          for (int which = 0; which < 2; ++which) {
            const std::vector<double>& src = which ? d.policy : d.regret;
            auto& ps = partial[{d.key, which == 1}];
            if (ps.empty()) ps.assign(64, std::vector<double>(src.size(), 0.0));
            for (size_t a = 0; a < src.size(); ++a) if (src[a] != 0.0) ps[k % 64][a] += src[a];
          }
      for (auto& kv : partial) {
        auto& ps = kv.second;
        for (int s = 32; s >= 1; s >>= 1)
          for (int q = 0; q < s; ++q)
            for (size_t a = 0; a < ps[q].size(); ++a) ps[q][a] += ps[q + s][a];
        OsValues& v = table[kv.first.first];
        for (size_t a = 0; a < ps[0].size(); ++a) (kv.first.second ? v.cum_policy : v.regrets)[a] += ps[0][a];
      }
    }
    ++iteration;
  }
};

}  // namespace
}  // namespace oracle

extern "C" {

void* orc_osmccfr_new(void* game, uint64_t seed, int rng_mode, int trajectories_per_update, double epsilon) {
  auto* m = new oracle::OsMccfr;
  m->game = (oracle::Game*)game;
  m->n = m->game->info.num_players;
  m->seed = seed; m->rng_mode = rng_mode; m->K = trajectories_per_update; m->epsilon = epsilon;
  m->mt.seed((uint32_t)seed);
  return m;
}
void orc_osmccfr_free(void* m) { delete (oracle::OsMccfr*)m; }
int orc_osmccfr_iterate(void* m, int iters) {
  auto* s = (oracle::OsMccfr*)m;
  for (int i = 0; i < iters; ++i) s->RunIteration();
  return s->failed ? 1 : 0;
}
int orc_osmccfr_num_infosets(void* m) { return (int)((oracle::OsMccfr*)m)->table.size(); }
int orc_osmccfr_get(void* m, int k, char* key, int key_cap, int64_t* legal, double* regrets, double* cum, int cap, int* player) {
  auto& table = ((oracle::OsMccfr*)m)->table;
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
