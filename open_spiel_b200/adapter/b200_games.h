// C++ host adapter: open_spiel::Game / State subclasses whose every rule computation is a call into the
// b2s C ABI (include/b2s.h -> libb2s.so -> sm_100a kernels).  Registered under the reference's own short names
// (GameRegisterer::RegisterGame overwrites, spiel.cc:216-219), so open_spiel::LoadGame("connect_four") returns
// these objects and everything written against State (algorithms, tests/basic_tests.cc, pyspiel) keeps working.
// A scalar State holds a ONE-lane device batch: this is the compatibility path; throughput comes from
// B200Game::NewBatch (raw b2s batch handle for vectorised callers).
#ifndef OPEN_SPIEL_B200_ADAPTER_B200_GAMES_H_
#define OPEN_SPIEL_B200_ADAPTER_B200_GAMES_H_

#include <memory>
#include <string>
#include <vector>

#include "open_spiel/spiel.h"

extern "C" {
#include "b2s.h"
}

namespace open_spiel {
namespace b200 {

class B200Game : public Game {
 public:
  B200Game(const GameType& type, const GameParameters& params);
  int NumDistinctActions() const override { return info_.num_distinct_actions; }
  std::unique_ptr<State> NewInitialState() const override;
  int NumPlayers() const override { return info_.num_players; }
  double MinUtility() const override { return info_.min_utility; }
  double MaxUtility() const override { return info_.max_utility; }
  absl::optional<double> UtilitySum() const override { return 0; }
  int MaxGameLength() const override { return info_.max_game_length; }
  std::vector<int> ObservationTensorShape() const override;
  // Vectorised entry point: a raw b2s batch of n lanes of this game (caller owns it; b2s_batch_destroy).
  void* NewBatch(int64_t n, int device = 0) const;
  int gid() const { return gid_; }
  const b2s_params& cparams() const { return cparams_; }
  const b2s_game_info& info() const { return info_; }

 private:
  int gid_;
  b2s_params cparams_;
  b2s_game_info info_;
};

class B200State : public State {
 public:
  explicit B200State(std::shared_ptr<const Game> game);
  B200State(const B200State& other);
  ~B200State() override;
  Player CurrentPlayer() const override;
  std::vector<Action> LegalActions() const override;
  std::string ActionToString(Player player, Action action_id) const override;
  std::string ToString() const override;
  bool IsTerminal() const override;
  std::vector<double> Returns() const override;
  std::string InformationStateString(Player player) const override { return HistoryString(); }
  std::string ObservationString(Player player) const override { return ToString(); }
  void ObservationTensor(Player player, absl::Span<float> values) const override;
  std::unique_ptr<State> Clone() const override;

 protected:
  void DoApplyAction(Action action_id) override;

 private:
  const B200Game& bgame() const { return static_cast<const B200Game&>(*game_); }
  void* batch_ = nullptr;
  void* scratch_d_ = nullptr;      // device scratch for one action / status / mask / tensor
};

// Registers the B200 implementations over the stock tic_tac_toe and connect_four (call after static init).
void RegisterB200Games();

}  // namespace b200
}  // namespace open_spiel
#endif
