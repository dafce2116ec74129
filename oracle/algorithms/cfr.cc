// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/algorithms/cfr.{h,cc}: CFRSolverBase with alternating updates
// (EvaluateAndUpdatePolicy :263-282, ComputeCounterFactualRegret :331-408, ...ForActionProbs :443-469,
// CounterFactualReachProb :309-318, AllPlayersHaveZeroReachProb :471-479, ApplyRegretMatching :596-615,
// ApplyRegretMatchingPlusReset :683-691, CFRAveragePolicy :104-125) — recursive, string-keyed, one State clone
// per edge, exactly like the reference, over the oracle's State objects.
#include <cstring>
#include <map>
#include <string>

#include "../oracle.h"

namespace oracle {
namespace {

struct Values {                        // CFRInfoStateValues, cfr.h:42-98
  std::vector<int64_t> legal;
  std::vector<double> regrets, cum_policy, cur_policy;
  int player = 0;
};

struct Cfr {
  const Game* game;
  bool linear = false, rm_plus = false;
  int iteration = 0;
  int n;                               // players
  std::map<std::string, Values> table;
  std::vector<std::string> order;      // first-visit (DFS) order of information states

  void Init(const State& s) {          // InitializeInfostateNodes, cfr.cc:234-261
    if (s.IsTerminal()) return;
    if (s.IsChanceNode()) {
      for (auto& ap : s.ChanceOutcomes()) { auto c = s.Clone(); c->ApplyAction(ap.first); Init(*c); }
      return;
    }
    int p = s.CurrentPlayer();
    std::string key = s.InformationStateString(p);
    auto la = s.LegalActions();
    if (!table.count(key)) order.push_back(key);
    Values v;
    v.legal = la; v.player = p;
    v.regrets.assign(la.size(), 0.0); v.cum_policy.assign(la.size(), 0.0);
    v.cur_policy.assign(la.size(), 1.0 / la.size());
    table[key] = v;
    for (auto a : la) { auto c = s.Clone(); c->ApplyAction(a); Init(*c); }
  }
  std::vector<double> ForProbs(const State& s, int upd, const std::vector<double>& reach, int cur,
                               const std::vector<double>& probs, const std::vector<int64_t>& acts,
                               std::vector<double>* child_out) {
    std::vector<double> value(n);
    for (size_t i = 0; i < acts.size(); ++i) {
      double prob = probs[i];
      auto ns = s.Clone();
      ns->ApplyAction(acts[i]);
      std::vector<double> nr(reach);
      nr[cur] *= prob;
      auto cv = Regret(*ns, upd, nr);
      for (int k = 0; k < n; ++k) value[k] += prob * cv[k];
      if (child_out) child_out->push_back(cv[cur]);
    }
    return value;
  }
  std::vector<double> Regret(const State& s, int upd, const std::vector<double>& reach) {
    if (s.IsTerminal()) return s.Returns();
    if (s.IsChanceNode()) {
      auto co = s.ChanceOutcomes();
      std::vector<double> dist; std::vector<int64_t> outs;
      for (auto& ap : co) { outs.push_back(ap.first); dist.push_back(ap.second); }
      return ForProbs(s, upd, reach, n, dist, outs, nullptr);
    }
    bool all_zero = true;
    for (int i = 0; i < n; ++i) if (reach[i] != 0.0) all_zero = false;
    if (all_zero) return std::vector<double>(n, 0.0);
    int cur = s.CurrentPlayer();
    std::string key = s.InformationStateString(cur);
    auto la = s.LegalActions();
    std::vector<double> policy = table[key].cur_policy;
    std::vector<double> child;
    auto value = ForProbs(s, upd, reach, cur, policy, la, &child);
    if (upd == cur) {
      Values v = table[key];
      double self = reach[cur], cfr = 1.0;
      for (size_t i = 0; i < reach.size(); ++i) if ((int)i != cur) cfr *= reach[i];
      for (size_t a = 0; a < la.size(); ++a) {
        double r = cfr * (child[a] - value[cur]);
        v.regrets[a] += r;
        if (linear) v.cum_policy[a] += iteration * self * policy[a];
        else v.cum_policy[a] += self * policy[a];
      }
      table[key] = v;
    }
    return value;
  }
  void Match() {
    for (auto& kv : table) {
      Values& v = kv.second;
      double sum = 0.0;
      for (double r : v.regrets) if (r > 0) sum += r;
      for (size_t a = 0; a < v.regrets.size(); ++a)
        v.cur_policy[a] = sum > 0 ? (v.regrets[a] > 0 ? v.regrets[a] / sum : 0) : 1.0 / v.legal.size();
    }
  }
  void Iterate() {
    ++iteration;
    auto root = game->NewInitialState();
    std::vector<double> reach(n + 1, 1.0);
    for (int p = 0; p < n; ++p) {
      Regret(*root, p, reach);
      if (rm_plus) for (auto& kv : table) for (double& r : kv.second.regrets) if (r < 0) r = 0;
      Match();
    }
  }
};

}  // namespace
}  // namespace oracle

extern "C" {

void* orc_cfr_new(void* game, int linear_averaging, int rm_plus) {
  using namespace oracle;
  auto* c = new Cfr;
  c->game = (Game*)game;
  c->n = c->game->info.num_players;
  c->linear = linear_averaging != 0;
  c->rm_plus = rm_plus != 0;
  auto root = c->game->NewInitialState();
  c->Init(*root);
  return c;
}
void orc_cfr_free(void* c) { delete (oracle::Cfr*)c; }
void orc_cfr_iterate(void* c, int iters) { for (int i = 0; i < iters; ++i) ((oracle::Cfr*)c)->Iterate(); }
int orc_cfr_num_infosets(void* c) { return (int)((oracle::Cfr*)c)->table.size(); }
// k-th information state in first-visit order: key string + arrays; returns the number of legal actions.
int orc_cfr_get(void* c, int k, char* key, int key_cap, int64_t* legal, double* regrets, double* cum, double* cur, int cap, int* player) {
  auto* s = (oracle::Cfr*)c;
  if (k < 0 || k >= (int)s->order.size()) return -1;
  const std::string& ks = s->order[k];
  const auto& v = s->table[ks];
  int m = (int)ks.size() < key_cap - 1 ? (int)ks.size() : key_cap - 1;
  memcpy(key, ks.data(), m); key[m] = 0;
  int n = (int)v.legal.size();
  for (int i = 0; i < n && i < cap; ++i) { legal[i] = v.legal[i]; regrets[i] = v.regrets[i]; cum[i] = v.cum_policy[i]; cur[i] = v.cur_policy[i]; }
  if (player) *player = v.player;
  return n;
}

}  // extern "C"
