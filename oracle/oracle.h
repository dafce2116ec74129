// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's per-game State
// transition functions and of the MCTS / CFR loops that drive them.
//
// Nothing in the product path (open_spiel_b200/, include/) may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
// and only as the checker.  Each game file cites the reference file:line it restates.
// The restatement deliberately keeps the reference's *array based* algorithms (cell arrays, full
// board scans, linked-list chains, recursion) so that it is independent of the bitboard kernels it
// checks.  Pinned against: tests/golden/playthroughs/*.json (reference playthrough traces), the
// known-answer cases of the reference's *_test.cc files (tests/test_oracle_known_answers.py), and,
// when built, the unmodified reference compiled against an abseil shim (oracle/_ref): tests/test_ref_vs_oracle.py (State
// functions, CFR tables), test_mcts_oracle_vs_reference.py (MCTSBot searches bit for bit on the reference's RNG streams),
// test_mccfr_oracle.py (ExternalSamplingMCCFRSolver tables bit for bit), test_trajectories_oracle.py (RecordBatchedTrajectory).
#ifndef B2S_ORACLE_H_
#define B2S_ORACLE_H_

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace oracle {

// Sentinels: reference open_spiel/spiel_globals.h:26-56,82.
constexpr int kChancePlayerId = -1;
constexpr int kTerminalPlayerId = -4;
constexpr int kInvalidPlayer = -3;
constexpr int64_t kInvalidAction = -1;

struct GameInfo {
  std::string name;
  int num_players = 2;
  int num_distinct_actions = 0;
  int max_game_length = 0;
  int max_chance_outcomes = 0;
  int observation_tensor_size = 0;
  int information_state_tensor_size = 0;
  double min_utility = -1, max_utility = 1;
};

struct Params {           // integer / double game parameters, by name
  std::vector<std::pair<std::string, double>> kv;
  double get(const std::string& k, double dflt) const {
    for (auto& p : kv) if (p.first == k) return p.second;
    return dflt;
  }
};

class State {
 public:
  virtual ~State() = default;
  virtual int CurrentPlayer() const = 0;
  virtual std::vector<int64_t> LegalActions() const = 0;
  virtual bool IsTerminal() const = 0;
  virtual std::vector<double> Returns() const = 0;
  virtual std::string ToString() const = 0;
  virtual std::unique_ptr<State> Clone() const = 0;
  virtual void ObservationTensor(int player, float* out) const = 0;
  virtual void InformationStateTensor(int, float*) const {}
  virtual std::string InformationStateString(int) const { return ""; }
  virtual std::string ObservationString(int) const { return ToString(); }
  virtual std::vector<std::pair<int64_t, double>> ChanceOutcomes() const { return {}; }
  // Superset of LegalActions() from which random playouts draw by rejection (a uniformly drawn candidate is kept
  // iff it is legal, which is a uniform draw over the legal actions — what mcts.cc:51-55 asks for).  Default: the
  // legal actions themselves (never rejected).  go overrides it with "empty points except the ko point, then
  // pass", which spares the device the full 81-point legality scan on every playout ply.
  virtual std::vector<int64_t> RolloutCandidates() const { return LegalActions(); }
  bool IsChanceNode() const { return CurrentPlayer() == kChancePlayerId; }

  // reference State::ApplyAction, open_spiel/spiel.cc:441-451
  void ApplyAction(int64_t a) {
    int p = CurrentPlayer();
    DoApplyAction(a);
    history_.push_back({p, a});
  }
  const std::vector<std::pair<int, int64_t>>& History() const { return history_; }
  bool error = false;      // set instead of SPIEL_CHECK-aborting
  std::string error_msg;

 protected:
  virtual void DoApplyAction(int64_t a) = 0;
  void Fail(const std::string& m) { error = true; error_msg = m; }
  std::vector<std::pair<int, int64_t>> history_;
};

class Game {
 public:
  virtual ~Game() = default;
  virtual std::unique_ptr<State> NewInitialState() const = 0;
  GameInfo info;
};

std::unique_ptr<Game> LoadGame(const std::string& name, const Params& params);

// per-game factories
std::unique_ptr<Game> MakeTicTacToe(const Params&);
std::unique_ptr<Game> MakeConnectFour(const Params&);
std::unique_ptr<Game> MakeBreakthrough(const Params&);
std::unique_ptr<Game> MakeHex(const Params&);
std::unique_ptr<Game> MakeGo(const Params&);
std::unique_ptr<Game> MakeKuhnPoker(const Params&);
std::unique_ptr<Game> MakeLeducPoker(const Params&);
std::unique_ptr<Game> MakeMnk(const Params&);
std::unique_ptr<Game> MakeOthello(const Params&);
std::unique_ptr<Game> MakeY(const Params&);
std::unique_ptr<Game> MakeHavannah(const Params&);

}  // namespace oracle
#endif  // B2S_ORACLE_H_
