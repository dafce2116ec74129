// TEST INFRASTRUCTURE ONLY.  Our own C ABI over the UNMODIFIED reference (open_spiel::Game / State /
// MCTSBot / CFRSolver), built by oracle/ref_build.mk into oracle/_ref/libspiel_ref_c.so.  It mirrors the
// orc_* surface of oracle/oracle_c.cc so tests can drive the restatement and the real reference alike.
#include <cstring>
#include <memory>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

#include "open_spiel/algorithms/cfr.h"
#include "open_spiel/algorithms/external_sampling_mccfr.h"
#include "open_spiel/algorithms/mcts.h"
#include "open_spiel/algorithms/outcome_sampling_mccfr.h"
#include "open_spiel/algorithms/tabular_exploitability.h"
#include "open_spiel/algorithms/trajectories.h"
#include "open_spiel/policy.h"
#include "open_spiel/spiel.h"
#include "open_spiel/spiel_utils.h"

using open_spiel::Action;
using open_spiel::Game;
using open_spiel::State;

namespace {
struct GameHolder { std::shared_ptr<const Game> game; };
thread_local std::string g_err;
void ThrowingHandler(const std::string& msg) { throw std::runtime_error(msg); }
struct Init { Init() { open_spiel::SetErrorHandler(ThrowingHandler); } } g_init;

int CopyStr(const std::string& s, char* buf, int cap) {
  int n = (int)s.size();
  if (buf && cap > 0) { int m = n < cap - 1 ? n : cap - 1; memcpy(buf, s.data(), m); buf[m] = 0; }
  return n;
}
}  // namespace

#define GUARD(stmt, onerr) try { stmt; } catch (const std::exception& e) { g_err = e.what(); onerr; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_load_game(const char* game_string) {
  GUARD(return new GameHolder{open_spiel::LoadGame(std::string(game_string))}, return nullptr);
}
void ref_free_game(void* g) { delete (GameHolder*)g; }
int ref_num_distinct_actions(void* g) { return ((GameHolder*)g)->game->NumDistinctActions(); }
int ref_num_players(void* g) { return ((GameHolder*)g)->game->NumPlayers(); }
int ref_max_game_length(void* g) { return ((GameHolder*)g)->game->MaxGameLength(); }
int ref_max_chance_outcomes(void* g) { return ((GameHolder*)g)->game->MaxChanceOutcomes(); }
int ref_observation_tensor_size(void* g) { GUARD(return ((GameHolder*)g)->game->ObservationTensorSize(), return 0); }
int ref_information_state_tensor_size(void* g) {
  auto& game = ((GameHolder*)g)->game;
  if (!game->GetType().provides_information_state_tensor) return 0;
  GUARD(return game->InformationStateTensorSize(), return 0);
}
double ref_min_utility(void* g) { return ((GameHolder*)g)->game->MinUtility(); }
double ref_max_utility(void* g) { return ((GameHolder*)g)->game->MaxUtility(); }

void* ref_new_initial_state(void* g) { GUARD(return ((GameHolder*)g)->game->NewInitialState().release(), return nullptr); }
void* ref_clone(void* s) { return ((State*)s)->Clone().release(); }
void ref_free_state(void* s) { delete (State*)s; }
int ref_current_player(void* s) { return ((State*)s)->CurrentPlayer(); }
int ref_is_terminal(void* s) { return ((State*)s)->IsTerminal() ? 1 : 0; }
int ref_legal_actions(void* s, int64_t* out, int cap) {
  std::vector<Action> v;
  GUARD(v = ((State*)s)->LegalActions(), return -1);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
int ref_apply_action(void* s, int64_t a) { GUARD(((State*)s)->ApplyAction(a); return 0, return 1); }
void ref_returns(void* s, double* out) {
  auto v = ((State*)s)->Returns();
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
int ref_observation_tensor(void* s, int player, float* out, int n) {
  GUARD(((State*)s)->ObservationTensor(player, absl::MakeSpan(out, n)); return 0, return 1);
}
int ref_information_state_tensor(void* s, int player, float* out, int n) {
  GUARD(((State*)s)->InformationStateTensor(player, absl::MakeSpan(out, n)); return 0, return 1);
}
int ref_to_string(void* s, char* buf, int cap) { GUARD(return CopyStr(((State*)s)->ToString(), buf, cap), return -1); }
int ref_information_state_string(void* s, int player, char* buf, int cap) {
  GUARD(return CopyStr(((State*)s)->InformationStateString(player), buf, cap), return -1);
}
int ref_observation_string(void* s, int player, char* buf, int cap) {
  GUARD(return CopyStr(((State*)s)->ObservationString(player), buf, cap), return -1);
}
int ref_chance_outcomes(void* s, int64_t* actions, double* probs, int cap) {
  std::vector<std::pair<Action, double>> v;
  GUARD(v = ((State*)s)->ChanceOutcomes(), return -1);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) { actions[i] = v[i].first; probs[i] = v[i].second; }
  return (int)v.size();
}
int ref_history(void* s, int64_t* out, int cap) {
  auto h = ((State*)s)->History();
  for (int i = 0; i < (int)h.size() && i < cap; ++i) out[i] = h[i];
  return (int)h.size();
}

// ---- CFRSolver (algorithms/cfr.h:312-328) -----------------------------------------------------------------
void* ref_cfr_new(void* g) { GUARD(return new open_spiel::algorithms::CFRSolver(*((GameHolder*)g)->game), return nullptr); }
void ref_cfr_free(void* c) { delete (open_spiel::algorithms::CFRSolver*)c; }
int ref_cfr_iterate(void* c, int iters) {
  GUARD(for (int i = 0; i < iters; ++i) ((open_spiel::algorithms::CFRSolver*)c)->EvaluateAndUpdatePolicy(); return 0, return 1);
}
int ref_cfr_num_infostates(void* c) { return (int)((open_spiel::algorithms::CFRSolver*)c)->InfoStateValuesTable().size(); }
// Table entry for an info-state key: copies up to cap values of each array; returns the number of legal actions, -1 if absent.
int ref_cfr_get(void* c, const char* key, int64_t* legal, double* regrets, double* cum_policy, double* cur_policy, int cap) {
  auto& table = ((open_spiel::algorithms::CFRSolver*)c)->InfoStateValuesTable();
  auto it = table.find(key);
  if (it == table.end()) return -1;
  const auto& v = it->second;
  int n = (int)v.legal_actions.size();
  for (int i = 0; i < n && i < cap; ++i) {
    legal[i] = v.legal_actions[i];
    regrets[i] = v.cumulative_regrets[i];
    cum_policy[i] = v.cumulative_policy[i];
    cur_policy[i] = v.current_policy[i];
  }
  return n;
}
// All keys, '\n'-separated (sorted).
int ref_cfr_keys(void* c, char* buf, int cap) {
  auto& table = ((open_spiel::algorithms::CFRSolver*)c)->InfoStateValuesTable();
  std::vector<std::string> keys;
  for (auto& kv : table) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  std::string s;
  for (auto& k : keys) { s += k; s += '\n'; }
  return CopyStr(s, buf, cap);
}
double ref_cfr_exploitability(void* g, void* c) {
  auto* solver = (open_spiel::algorithms::CFRSolver*)c;
  GUARD(return open_spiel::algorithms::Exploitability(*((GameHolder*)g)->game, *solver->AveragePolicy()), return -1.0);
}
double ref_cfr_nash_conv(void* g, void* c) {
  auto* solver = (open_spiel::algorithms::CFRSolver*)c;
  GUARD(return open_spiel::algorithms::NashConv(*((GameHolder*)g)->game, *solver->AveragePolicy()), return -1.0);
}

// ---- MCTSBot (algorithms/mcts.h:149-230) --------------------------------------------------------------------
// Runs MCTSearch from `state` and reports the root's children (action, visits, total reward, proven outcome flag).
int ref_mcts_search_mb(void* g, void* state, double uct_c, int max_simulations, int n_rollouts, int solve, int seed,
                       int64_t* child_actions, int* child_visits, double* child_rewards, int cap, int64_t* best_action,
                       int* root_visits, int max_memory_mb);
int ref_mcts_search(void* g, void* state, double uct_c, int max_simulations, int n_rollouts, int solve, int seed,
                    int64_t* child_actions, int* child_visits, double* child_rewards, int cap, int64_t* best_action,
                    int* root_visits) {
  return ref_mcts_search_mb(g, state, uct_c, max_simulations, n_rollouts, solve, seed, child_actions, child_visits, child_rewards, cap,
                            best_action, root_visits, 1000);
}
// sizeof(SearchNode) decides MCTSBot::max_nodes_ = (max_memory_mb << 20) / sizeof(SearchNode) + 1 (mcts.cc:214)
int ref_sizeof_search_node() { return (int)sizeof(open_spiel::algorithms::SearchNode); }
int ref_mcts_search_mb(void* g, void* state, double uct_c, int max_simulations, int n_rollouts, int solve, int seed,
                       int64_t* child_actions, int* child_visits, double* child_rewards, int cap, int64_t* best_action,
                       int* root_visits, int max_memory_mb) {
  try {
    auto game = ((GameHolder*)g)->game;
    auto evaluator = std::make_shared<open_spiel::algorithms::RandomRolloutEvaluator>(n_rollouts, seed);
    open_spiel::algorithms::MCTSBot bot(*game, evaluator, uct_c, max_simulations, max_memory_mb, solve != 0, seed,
                                        /*verbose=*/false);
    std::unique_ptr<open_spiel::algorithms::SearchNode> root = bot.MCTSearch(*(State*)state);
    int n = (int)root->children.size();
    for (int i = 0; i < n && i < cap; ++i) {
      child_actions[i] = root->children[i].action;
      child_visits[i] = root->children[i].explore_count;
      child_rewards[i] = root->children[i].total_reward;
    }
    if (best_action) *best_action = root->BestChild().action;
    if (root_visits) *root_visits = root->explore_count;
    return n;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- RecordBatchedTrajectory (algorithms/trajectories.h:86-100) with GetUniformPolicy for every player -----------
// Flattens the reference's BatchedTrajectory: observations [B][T][F], legal [B][T][A], policies [B][T][A], actions /
// players / valid / next_is_terminal [B][T], rewards [B][P].  Returns max_trajectory_length (== T after ResizeFields).
int ref_record_batched_trajectory(void* g, int batch_size, int seed, int T, float* observations, int* legal,
                                  double* policies, int64_t* actions, int* players, int* valid, int* next_is_terminal,
                                  double* rewards) {
  try {
    auto game = ((GameHolder*)g)->game;
    std::vector<open_spiel::TabularPolicy> pol(game->NumPlayers(), open_spiel::GetUniformPolicy(*game));
    open_spiel::algorithms::BatchedTrajectory bt = open_spiel::algorithms::RecordBatchedTrajectory(
        *game, pol, /*state_to_index=*/{}, batch_size, /*include_full_observations=*/true, seed, T);
    const int A = game->NumDistinctActions(), P = game->NumPlayers();
    const int len = (int)bt.max_trajectory_length;
    for (int b = 0; b < batch_size; ++b) {
      for (int t = 0; t < len; ++t) {
        const auto& o = bt.observations[b][t];
        const size_t F = o.size();
        for (size_t f = 0; f < F; ++f) observations[((size_t)b * len + t) * F + f] = o[f];
        for (int a = 0; a < A; ++a) {
          legal[((size_t)b * len + t) * A + a] = bt.legal_actions[b][t][a];
          policies[((size_t)b * len + t) * A + a] = bt.player_policies[b][t][a];
        }
        actions[(size_t)b * len + t] = bt.actions[b][t];
        players[(size_t)b * len + t] = bt.player_ids[b][t];
        valid[(size_t)b * len + t] = bt.valid[b][t];
        next_is_terminal[(size_t)b * len + t] = bt.next_is_terminal[b][t];
      }
      for (int p = 0; p < P; ++p) rewards[(size_t)b * P + p] = bt.rewards[b][p];
    }
    return len;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- ExternalSamplingMCCFRSolver (algorithms/external_sampling_mccfr.h:55-110), AverageType::kSimple ----------
void* ref_mccfr_new(void* g, int seed) {
  GUARD(return new open_spiel::algorithms::ExternalSamplingMCCFRSolver(*((GameHolder*)g)->game, seed), return nullptr);
}
void* ref_mccfr_new_full(void* g, int seed) {      // AverageType::kFull (external_sampling_mccfr.h:53-54)
  GUARD(return new open_spiel::algorithms::ExternalSamplingMCCFRSolver(*((GameHolder*)g)->game, seed,
                                                                        open_spiel::algorithms::AverageType::kFull), return nullptr);
}
void ref_mccfr_free(void* c) { delete (open_spiel::algorithms::ExternalSamplingMCCFRSolver*)c; }
int ref_mccfr_iterate(void* c, int iters) {
  GUARD(for (int i = 0; i < iters; ++i) ((open_spiel::algorithms::ExternalSamplingMCCFRSolver*)c)->RunIteration(); return 0,
        return 1);
}
int ref_mccfr_keys(void* c, char* buf, int cap) {
  auto& table = ((open_spiel::algorithms::ExternalSamplingMCCFRSolver*)c)->InfoStateValuesTable();
  std::vector<std::string> keys;
  for (auto& kv : table) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  std::string s;
  for (auto& k : keys) { s += k; s += '\n'; }
  return CopyStr(s, buf, cap);
}
int ref_mccfr_get(void* c, const char* key, int64_t* legal, double* regrets, double* cum_policy, int cap) {
  auto& table = ((open_spiel::algorithms::ExternalSamplingMCCFRSolver*)c)->InfoStateValuesTable();
  auto it = table.find(key);
  if (it == table.end()) return -1;
  const auto& v = it->second;
  int n = (int)v.legal_actions.size();
  for (int i = 0; i < n && i < cap; ++i) {
    legal[i] = v.legal_actions[i];
    regrets[i] = v.cumulative_regrets[i];
    cum_policy[i] = v.cumulative_policy[i];
  }
  return n;
}
double ref_mccfr_nash_conv(void* g, void* c) {
  auto* solver = (open_spiel::algorithms::ExternalSamplingMCCFRSolver*)c;
  GUARD(return open_spiel::algorithms::NashConv(*((GameHolder*)g)->game, *solver->AveragePolicy(), /*use_state_get_policy=*/true),
               return -1.0);
}

// ---- OutcomeSamplingMCCFRSolver (algorithms/outcome_sampling_mccfr.h:40-66) -------------------------------------------
using OsSolver = open_spiel::algorithms::OutcomeSamplingMCCFRSolver;
void* ref_osmccfr_new(void* g, double epsilon, int seed) { GUARD(return new OsSolver(*((GameHolder*)g)->game, epsilon, seed), return nullptr); }
void ref_osmccfr_free(void* c) { delete (OsSolver*)c; }
int ref_osmccfr_iterate(void* c, int iters) { GUARD(for (int i = 0; i < iters; ++i) ((OsSolver*)c)->RunIteration(); return 0, return 1); }
int ref_osmccfr_keys(void* c, char* buf, int cap) {
  std::vector<std::string> keys;
  for (auto& kv : ((OsSolver*)c)->InfoStateValuesTable()) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  std::string s;
  for (auto& k : keys) { s += k; s += '\n'; }
  return CopyStr(s, buf, cap);
}
int ref_osmccfr_get(void* c, const char* key, int64_t* legal, double* regrets, double* cum_policy, int cap) {
  auto& table = ((OsSolver*)c)->InfoStateValuesTable();
  auto it = table.find(key);
  if (it == table.end()) return -1;
  const auto& v = it->second;
  int n = (int)v.legal_actions.size();
  for (int i = 0; i < n && i < cap; ++i) { legal[i] = v.legal_actions[i]; regrets[i] = v.cumulative_regrets[i]; cum_policy[i] = v.cumulative_policy[i]; }
  return n;
}
double ref_osmccfr_nash_conv(void* g, void* c) {
  GUARD(return open_spiel::algorithms::NashConv(*((GameHolder*)g)->game, *((OsSolver*)c)->AveragePolicy(), /*use_state_get_policy=*/true), return -1.0);
}

// ---- text formats (Game::ToString spiel.cc:802-806, CFRSolverBase::Serialize cfr.cc:284-307, DeserializeCFRSolver :704-715) ----
int ref_game_to_string(void* g, char* buf, int cap) { GUARD(return CopyStr(((GameHolder*)g)->game->ToString(), buf, cap), return -1); }
int ref_cfr_serialize(void* c, char* buf, int cap) {
  GUARD(return CopyStr(((open_spiel::algorithms::CFRSolver*)c)->Serialize(), buf, cap), return -1);
}
void* ref_cfr_deserialize(const char* text) {
  GUARD(return open_spiel::algorithms::DeserializeCFRSolver(text).release(), return nullptr);
}
int ref_state_serialize(void* s, char* buf, int cap) { GUARD(return CopyStr(((State*)s)->Serialize(), buf, cap), return -1); }
void* ref_deserialize_state(void* g, const char* text) {
  GUARD(return ((GameHolder*)g)->game->DeserializeState(text).release(), return nullptr);
}


// Replays n recorded lanes on the unmodified reference, `threads` at a time: lane i applies hist[i*L .. i*L+L) (entries
// < 0 are skipped) from the initial state, then `final_action[i]` (if >= 0).  Outputs, any nullable:
//   mask_before [n][W] u32  LegalActions() as bits just before the final action
//   terminal [n] u8, cur_player [n] i8, returns [n][P] f32, mask_after [n][W] u32 after it
//   obs_bits [n][OW] u32    ObservationTensor(0) != 0 as a bit string (OW = ceil(size / 32))
// Returns the number of lanes on which the reference raised an error (illegal action), -1 on setup failure.
long ref_replay_batch(void* g, long n, int L, const int32_t* hist, const int32_t* final_action, int W, int OW, int threads,
                      uint32_t* mask_before, uint8_t* terminal, int8_t* cur_player, float* returns, uint32_t* mask_after,
                      uint32_t* obs_bits) {
  auto game = ((GameHolder*)g)->game;
  const int P = game->NumPlayers();
  int obs_size = 0;
  GUARD(obs_size = game->ObservationTensorSize(), return -1);
  if (threads < 1) threads = 1;
  std::vector<long> bad(threads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      std::vector<float> obs(obs_size);
      for (long i = n * t / threads; i < n * (t + 1) / threads; ++i) {
        try {
          std::unique_ptr<State> s = game->NewInitialState();
          for (int k = 0; k < L; ++k) if (hist[i * L + k] >= 0) s->ApplyAction(hist[i * L + k]);
          auto bits = [&](uint32_t* out) {
            for (int w = 0; w < W; ++w) out[w] = 0;
            for (Action a : s->LegalActions()) out[a >> 5] |= 1u << (a & 31);
          };
          if (mask_before) bits(mask_before + i * W);
          if (final_action[i] >= 0) s->ApplyAction(final_action[i]);
          if (terminal) terminal[i] = s->IsTerminal() ? 1 : 0;
          if (cur_player) cur_player[i] = (int8_t)s->CurrentPlayer();
          if (returns) { auto r = s->Returns(); for (int p = 0; p < P; ++p) returns[i * P + p] = (float)r[p]; }
          if (mask_after) bits(mask_after + i * W);
          if (obs_bits) {
            s->ObservationTensor(0, absl::MakeSpan(obs));
            for (int w = 0; w < OW; ++w) obs_bits[i * OW + w] = 0;
            for (int e = 0; e < obs_size; ++e) if (obs[e] != 0.f) obs_bits[i * OW + (e >> 5)] |= 1u << (e & 31);
          }
        } catch (const std::exception&) {
          ++bad[t];
        }
      }
    });
  }
  for (auto& th : pool) th.join();
  long total = 0;
  for (long b : bad) total += b;
  return total;
}
}  // extern "C"
