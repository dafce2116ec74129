// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  extern "C" surface used by tests via ctypes.
#include <cstring>
#include <map>

#include "oracle.h"

namespace oracle {
std::unique_ptr<Game> LoadGame(const std::string& name, const Params& p) {
  if (name == "tic_tac_toe") return MakeTicTacToe(p);
  if (name == "connect_four") return MakeConnectFour(p);
  if (name == "breakthrough") return MakeBreakthrough(p);
  if (name == "hex") return MakeHex(p);
  if (name == "go") return MakeGo(p);
  if (name == "kuhn_poker") return MakeKuhnPoker(p);
  if (name == "leduc_poker") return MakeLeducPoker(p);
  if (name == "mnk") return MakeMnk(p);
  if (name == "othello") return MakeOthello(p);
  if (name == "y") return MakeY(p);
  if (name == "havannah") return MakeHavannah(p);
  return nullptr;
}
}  // namespace oracle

using oracle::Game;
using oracle::State;

static int CopyStr(const std::string& s, char* buf, int cap) {
  int n = (int)s.size();
  if (buf && cap > 0) {
    int m = n < cap - 1 ? n : cap - 1;
    memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return n;
}

extern "C" {

// params: parallel arrays of names / values.
void* orc_load_game(const char* name, int n_params, const char** keys, const double* vals) {
  oracle::Params p;
  for (int i = 0; i < n_params; ++i) p.kv.push_back({keys[i], vals[i]});
  auto g = oracle::LoadGame(name, p);
  return g.release();
}
void orc_free_game(void* g) { delete (Game*)g; }
int orc_num_distinct_actions(void* g) { return ((Game*)g)->info.num_distinct_actions; }
int orc_num_players(void* g) { return ((Game*)g)->info.num_players; }
int orc_max_game_length(void* g) { return ((Game*)g)->info.max_game_length; }
int orc_observation_tensor_size(void* g) { return ((Game*)g)->info.observation_tensor_size; }
int orc_information_state_tensor_size(void* g) { return ((Game*)g)->info.information_state_tensor_size; }
int orc_max_chance_outcomes(void* g) { return ((Game*)g)->info.max_chance_outcomes; }
double orc_min_utility(void* g) { return ((Game*)g)->info.min_utility; }
double orc_max_utility(void* g) { return ((Game*)g)->info.max_utility; }

void* orc_new_initial_state(void* g) { return ((Game*)g)->NewInitialState().release(); }
void* orc_clone(void* s) { return ((State*)s)->Clone().release(); }
void orc_free_state(void* s) { delete (State*)s; }
int orc_current_player(void* s) { return ((State*)s)->CurrentPlayer(); }
int orc_is_terminal(void* s) { return ((State*)s)->IsTerminal() ? 1 : 0; }
int orc_legal_actions(void* s, int64_t* out, int cap) {
  auto v = ((State*)s)->LegalActions();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
int orc_rollout_candidates(void* s, int64_t* out, int cap) {
  auto v = ((State*)s)->RolloutCandidates();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
int orc_apply_action(void* s, int64_t a) {
  State* st = (State*)s;
  st->ApplyAction(a);
  return st->error ? 1 : 0;
}
void orc_returns(void* s, double* out) {
  auto v = ((State*)s)->Returns();
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void orc_observation_tensor(void* s, int player, float* out) { ((State*)s)->ObservationTensor(player, out); }
void orc_information_state_tensor(void* s, int player, float* out) { ((State*)s)->InformationStateTensor(player, out); }
int orc_to_string(void* s, char* buf, int cap) { return CopyStr(((State*)s)->ToString(), buf, cap); }
int orc_information_state_string(void* s, int player, char* buf, int cap) {
  return CopyStr(((State*)s)->InformationStateString(player), buf, cap);
}
int orc_observation_string(void* s, int player, char* buf, int cap) {
  return CopyStr(((State*)s)->ObservationString(player), buf, cap);
}
int orc_chance_outcomes(void* s, int64_t* actions, double* probs, int cap) {
  auto v = ((State*)s)->ChanceOutcomes();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) { actions[i] = v[i].first; probs[i] = v[i].second; }
  return (int)v.size();
}
int orc_history(void* s, int64_t* out, int cap) {
  auto& h = ((State*)s)->History();
  for (int i = 0; i < (int)h.size() && i < cap; ++i) out[i] = h[i].second;
  return (int)h.size();
}
int orc_error(void* s, char* buf, int cap) {
  State* st = (State*)s;
  if (!st->error) return 0;
  CopyStr(st->error_msg, buf, cap);
  return 1;
}

}  // extern "C"
