// C++ host adapter: open_spiel::Game / State subclasses for the seven b2s games, registered under the reference's own
// short names (GameRegisterer::RegisterGame overwrites, spiel.cc:216-219), so open_spiel::LoadGame("go(board_size=9)")
// and pyspiel.load_game(...) return these objects and everything written against State (algorithms,
// tests/basic_tests.cc, pyspiel) keeps working.
//
//  * A scalar B200State is a packed state blob (the b2s lane format) advanced by the HOST build of the same rule cores
//    the kernels are instantiated from (host_rules.h) — no kernel launch, no PCIe round trip per method call.
//  * Throughput comes from batches: B200Game::NewBatch returns a raw b2s batch handle (libb2s.so, GPU only), and
//    B200State::ToBatchLane / FromBatchLane move a scalar state into / out of a device lane unchanged.
//  * Parameter sets the packed layouts cannot hold (go > 9x9, 3-player poker, leduc action_mapping, ...) are served by
//    the stock reference game: the factory falls back to it, exactly as before registration.
#ifndef OPEN_SPIEL_B200_ADAPTER_B200_GAMES_H_
#define OPEN_SPIEL_B200_ADAPTER_B200_GAMES_H_

#include <memory>
#include <string>
#include <vector>

#include "host_rules.h"
#include "open_spiel/spiel.h"

namespace open_spiel {
namespace b200 {

class B200Game : public Game {
 public:
  // nullptr when the parameters are not representable (the caller falls back to the stock factory)
  static std::shared_ptr<const Game> Create(const GameType& type, const GameParameters& params);

  int NumDistinctActions() const override { TouchBoardParams(); return info().num_distinct_actions; }
  std::unique_ptr<State> NewInitialState() const override;
  int NumPlayers() const override { return info().num_players; }
  double MinUtility() const override { return info().min_utility; }
  double MaxUtility() const override { return info().max_utility; }
  absl::optional<double> UtilitySum() const override { return 0; }
  int MaxGameLength() const override { TouchBoardParams(); return info().max_game_length; }
  int MaxChanceOutcomes() const override { return info().max_chance_outcomes; }
  std::vector<int> ObservationTensorShape() const override;
  std::vector<int> InformationStateTensorShape() const override;
  std::string ActionToString(Player player, Action action_id) const override;
  // kuhn_poker / leduc_poker: the games' own structured observers (named tensor fields, public / private observation
  // strings), kuhn_poker.cc:428-437, leduc_poker.cc:853-862; the board games use the built-in observers.
  std::shared_ptr<Observer> MakeObserver(absl::optional<IIGObservationType> iig_obs_type,
                                         const GameParameters& params) const override;

  // Vectorised entry point: a raw b2s batch of n lanes of this game on `device` (caller owns it; b2s_batch_destroy).
  void* NewBatch(int64_t n, int device = 0) const;
  int gid() const { return gid_; }
  const b2s_params& cparams() const { return cparams_; }
  const b2s_game_info& info() const { return rules_->info(); }
  const b2s_host::Rules& rules() const { return *rules_; }
  bool hex_explicit() const { return hex_explicit_; }
  // The stock mnk reads its parameters only inside the methods that need them (mnk.h:120-123), and ParameterValue<>
  // records a defaulted parameter on its first read: GetParameters() / ToString() / serialization therefore print
  // "{m=15,n=15}" until the first move reads "k" (mnk.cc:167).  These two calls reproduce that from the same methods.
  void TouchBoardParams() const { if (gid_ == B2S_MNK) { ParameterValue<int>("n"); ParameterValue<int>("m"); } }
  void TouchLineParam() const { if (gid_ == B2S_MNK) ParameterValue<int>("k"); }
  float komi() const { return komi_; }

 private:
  B200Game(const GameType& type, const GameParameters& params) : Game(type, params) {}
  int gid_ = -1;
  b2s_params cparams_;
  std::unique_ptr<b2s_host::Rules> rules_;
  bool hex_explicit_ = false;    // hex string_rep=explicit (hex.cc:193-217)
  float komi_ = 7.5f;
};

class B200State : public State {
 public:
  explicit B200State(std::shared_ptr<const Game> game);
  B200State(const B200State&) = default;
  Player CurrentPlayer() const override;
  std::vector<Action> LegalActions() const override;
  std::string ActionToString(Player player, Action action_id) const override;
  std::string ToString() const override;
  bool IsTerminal() const override;
  std::vector<double> Returns() const override;
  std::string InformationStateString(Player player) const override;
  std::string ObservationString(Player player) const override;
  void ObservationTensor(Player player, absl::Span<float> values) const override;
  void InformationStateTensor(Player player, absl::Span<float> values) const override;
  std::unique_ptr<State> Clone() const override;
  void UndoAction(Player player, Action action) override;
  std::vector<std::pair<Action, double>> ChanceOutcomes() const override;

  // kuhn_poker / leduc_poker: what the poker observers and strings are made of (reference member names).
  struct PokerView {
    int num_players = 2;
    std::vector<int> private_cards;       // kInvalidCard (-10000) when not dealt (leduc), history-derived for kuhn
    int public_card = -10000, round = 1, cur_player = -1, pot = 0;
    std::vector<double> money;
    std::vector<int> ante;
    std::vector<int> round1, round2;      // leduc betting sequences: 0 fold, 1 call, 2 raise
  };
  PokerView Poker() const;

  // The packed lane (b2s_state_get / b2s_state_set layout).
  const void* blob() const { return blob_.data(); }
  size_t blob_bytes() const { return bgame().rules().blob_bytes(); }
  // Copies this state into lane `lane` of a b2s batch of the same game / from it.  FromBatchLane does not reconstruct
  // history_ (except for kuhn_poker, whose packed state is the history): use it on states whose lanes were advanced by
  // actions the caller also ApplyAction()s, or for read-only inspection of a device lane; UndoAction stops there.
  void ToBatchLane(void* batch, int64_t lane) const;
  void FromBatchLane(void* batch, int64_t lane);

 protected:
  void DoApplyAction(Action action_id) override;

 private:
  struct alignas(16) Word16 { uint64_t a, b; };
  const B200Game& bgame() const { return static_cast<const B200Game&>(*game_); }
  const b2s_host::Rules& rules() const { return bgame().rules(); }
  std::vector<Word16> blob_;           // packed state (+ go hash history)
  std::vector<Word16> undo_;           // stack of the state part of earlier blobs (history entries are never rewritten below ply)
};

// Registers the B200 implementations over the stock registrations of tic_tac_toe, connect_four, breakthrough, hex, go,
// kuhn_poker, leduc_poker, mnk, othello, y and havannah (call once, after static initialisation).  Idempotent.
void RegisterB200Games();

}  // namespace b200
}  // namespace open_spiel
#endif
