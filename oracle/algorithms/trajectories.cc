// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// Restatement of algorithms::RecordTrajectory / RecordBatchedTrajectory (reference
// open_spiel/algorithms/trajectories.cc:140-200 and :98-118) with uniform-random policies (GetUniformPolicy) and
// the injected Philox stream of the device recorder (open_spiel_b200/csrc/batch_kernels.cuh k_traj_step):
//   decision of step t:        k = PhiloxUniform(seed, lane, 64 (t+1), #legal)       -> k-th legal action
//   j-th chance node after it: k = PhiloxUniform(seed, lane, 64 (t+1) + 1 + j, #outcomes)   (before step 0: 1 + j)
// Episode-at-a-time like the reference; padding to T as BatchedTrajectory::ResizeFields (:62-96).
#include <cstdint>
#include <vector>

#include "../oracle.h"
#include "philox.h"

namespace oracle {
namespace {

// `forced` (optional): the complete action sequence of the episode, chance outcomes included; when given, actions
// are taken from it instead of being sampled (used to line an episode up with one recorded by the reference).
struct Forced {
  const int64_t* a = nullptr;
  int n = 0, k = 0;
  bool on() const { return a != nullptr; }
  int64_t next() { return k < n ? a[k++] : kInvalidAction; }
};

void ResolveChance(State* s, uint64_t seed, uint64_t lane, uint32_t b0, Forced* forced) {
  uint32_t j = 0;
  while (!s->IsTerminal() && s->IsChanceNode()) {
    auto outcomes = s->ChanceOutcomes();           // uniform in kuhn / leduc; listed in ascending action order
    if (forced->on()) {
      s->ApplyAction(forced->next());
    } else {
      uint32_t k = PhiloxUniform(seed, lane, b0 + 1u + j, (uint32_t)outcomes.size());
      s->ApplyAction(outcomes[k].first);
    }
    ++j;
  }
}

}  // namespace
}  // namespace oracle

extern "C" {

// One episode from a clone of `state`.  Outputs are [T]-padded rows of one batch entry: legal [T][A] ints (padding 1),
// observations [T][F] (padding 0; InformationStateTensor of the acting player if use_infostate, else
// ObservationTensor), actions / players / valid / next_is_terminal [T] (padding 0), rewards [P].
// Returns the episode length, or -1 if it did not finish within T decisions.
int orc_record_trajectory(void* game, void* state, uint64_t seed, uint64_t lane, int T, int use_infostate, int* legal,
                          float* observations, int64_t* actions, int* players, int* valid, int* next_is_terminal,
                          double* rewards, const int64_t* forced_actions, int n_forced) {
  using namespace oracle;
  Forced forced;
  forced.a = forced_actions; forced.n = n_forced;
  Game* g = (Game*)game;
  const int A = g->info.num_distinct_actions;
  const int F = use_infostate ? g->info.information_state_tensor_size : g->info.observation_tensor_size;
  auto s = ((State*)state)->Clone();
  for (int t = 0; t < T; ++t) {
    for (int a = 0; a < A; ++a) legal[t * A + a] = 1;
    if (observations) for (int f = 0; f < F; ++f) observations[(size_t)t * F + f] = 0.f;
    actions[t] = 0; players[t] = 0; valid[t] = 0; next_is_terminal[t] = 0;
  }
  ResolveChance(s.get(), seed, lane, 0u, &forced);
  int t = 0;
  while (!s->IsTerminal()) {
    if (t >= T) return -1;
    auto la = s->LegalActions();
    for (int a = 0; a < A; ++a) legal[t * A + a] = 0;
    for (auto a : la) legal[t * A + a] = 1;                       // State::LegalActionsMask
    int p = s->CurrentPlayer();
    if (observations) {
      if (use_infostate) s->InformationStateTensor(p, observations + (size_t)t * F);
      else s->ObservationTensor(p, observations + (size_t)t * F);
    }
    int64_t a = forced.on() ? forced.next() : la[PhiloxUniform(seed, lane, 64u * (uint32_t)(t + 1), (uint32_t)la.size())];
    players[t] = p;
    actions[t] = a;
    valid[t] = 1;
    s->ApplyAction(a);
    ResolveChance(s.get(), seed, lane, 64u * (uint32_t)(t + 1), &forced);
    ++t;
  }
  if (t > 0) next_is_terminal[t - 1] = 1;
  auto r = s->Returns();
  for (size_t i = 0; i < r.size(); ++i) rewards[i] = r[i];
  return t;
}

}  // extern "C"
