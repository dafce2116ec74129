// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/kuhn_poker/kuhn_poker.{h,cc} (n players, n+1 cards).
#include "../oracle.h"

namespace oracle {
namespace {

class KuhnState : public State {
 public:
  explicit KuhnState(int n) : n_(n), card_dealt_(n + 1, kInvalidPlayer), pot_(n), ante_(n, 1) {}  // :168-176

  // kuhn_poker.cc:178-185
  int CurrentPlayer() const override {
    if (IsTerminal()) return kTerminalPlayerId;
    return (int)history_.size() < n_ ? kChancePlayerId : (int)history_.size() % n_;
  }
  // kuhn_poker.cc:231-242
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    if (IsChanceNode()) {
      for (int c = 0; c < (int)card_dealt_.size(); ++c) if (card_dealt_[c] == kInvalidPlayer) v.push_back(c);
      return v;
    }
    return {0, 1};
  }
  bool IsTerminal() const override { return winner_ != kInvalidPlayer; }
  // kuhn_poker.cc:272-283
  std::vector<double> Returns() const override {
    std::vector<double> r(n_, 0.0);
    if (!IsTerminal()) return r;
    for (int p = 0; p < n_; ++p) {
      int bet = DidBet(p, history_) ? 2 : 1;
      r[p] = p == winner_ ? pot_ - bet : -bet;
    }
    return r;
  }
  // kuhn_poker.cc:254-268
  std::string ToString() const override {
    std::string s;
    for (int i = 0; i < (int)history_.size() && i < n_; ++i) {
      if (!s.empty()) s += ' ';
      s += std::to_string(history_[i].second);
    }
    if ((int)history_.size() > n_) s += ' ';
    for (int i = n_; i < (int)history_.size(); ++i) s += history_[i].second ? 'b' : 'p';
    return s;
  }
  // Observer, perfect recall (info state): kuhn_poker.cc:72-107, 109-166
  std::string InformationStateString(int player) const override {
    std::string s;
    if ((int)history_.size() > player) s += std::to_string(history_[player].second);
    for (int i = n_; i < (int)history_.size(); ++i) s += history_[i].second ? 'b' : 'p';
    return s;
  }
  std::string ObservationString(int player) const override {
    std::string s;
    if ((int)history_.size() > player) {
      s += std::to_string(history_[player].second);
      for (int p = 0; p < n_; ++p) s += std::to_string(ante_[p]);
    }
    return s;
  }
  void InformationStateTensor(int player, float* out) const override {
    int sz = 6 * n_ - 1;
    for (int i = 0; i < sz; ++i) out[i] = 0.f;
    out[player] = 1.f;                                               // "player" {n}
    if ((int)history_.size() > player) out[n_ + history_[player].second] = 1.f;   // "private_card" {n+1}
    float* bet = out + n_ + n_ + 1;                                  // "betting" {2n-1, 2}
    for (int i = n_; i < (int)history_.size(); ++i) bet[(i - n_) * 2 + history_[i].second] = 1.f;
  }
  void ObservationTensor(int player, float* out) const override {
    int sz = 3 * n_ + 1;
    for (int i = 0; i < sz; ++i) out[i] = 0.f;
    out[player] = 1.f;
    if ((int)history_.size() > player) out[n_ + history_[player].second] = 1.f;
    for (int p = 0; p < n_; ++p) out[2 * n_ + 1 + p] = (float)ante_[p];   // "pot_contribution" {n}
  }
  // kuhn_poker.cc:329-337
  std::vector<std::pair<int64_t, double>> ChanceOutcomes() const override {
    std::vector<std::pair<int64_t, double>> o;
    double p = 1.0 / (n_ + 1 - history_.size());
    for (int c = 0; c < (int)card_dealt_.size(); ++c) if (card_dealt_[c] == kInvalidPlayer) o.push_back({c, p});
    return o;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<KuhnState>(*this); }

 protected:
  // kuhn_poker.cc:190-229
  void DoApplyAction(int64_t move) override {
    if ((int)history_.size() < n_) {
      if (move < 0 || move > n_ || card_dealt_[move] != kInvalidPlayer) { Fail("kuhn: bad deal"); return; }
      card_dealt_[move] = (int)history_.size();
    } else if (move == 1) {
      if (first_bettor_ == kInvalidPlayer) first_bettor_ = CurrentPlayer();
      pot_ += 1;
      ante_[CurrentPlayer()] += 1;
    } else if (move != 0) { Fail("kuhn: bad action"); return; }
    auto h = history_;
    h.push_back({CurrentPlayer(), move});
    int num_actions = (int)h.size() - n_;
    if (first_bettor_ == kInvalidPlayer && num_actions == n_) {
      winner_ = card_dealt_[n_];
      if (winner_ == kInvalidPlayer) winner_ = card_dealt_[n_ - 1];
    } else if (first_bettor_ != kInvalidPlayer && num_actions == n_ + first_bettor_) {
      for (int card = n_; card >= 0; --card) {
        int p = card_dealt_[card];
        if (p != kInvalidPlayer && DidBet(p, h)) { winner_ = p; break; }
      }
    }
  }

 private:
  // kuhn_poker.cc:339-349
  bool DidBet(int player, const std::vector<std::pair<int, int64_t>>& h) const {
    if (first_bettor_ == kInvalidPlayer) return false;
    if (player == first_bettor_) return true;
    if (player > first_bettor_) return h[n_ + player].second == 1;
    return h[n_ * 2 + player].second == 1;
  }
  int n_;
  int first_bettor_ = kInvalidPlayer;
  std::vector<int> card_dealt_;
  int winner_ = kInvalidPlayer;
  int pot_;
  std::vector<int> ante_;
};

class KuhnGame : public Game {
 public:
  explicit KuhnGame(const Params& p) {
    n_ = (int)p.get("players", 2);
    info.name = "kuhn_poker";
    info.num_players = n_;
    info.num_distinct_actions = 2;                       // kuhn_poker.h:107
    info.max_chance_outcomes = n_ + 1;                   // kuhn_poker.h:111
    info.max_game_length = n_ * 2 - 1;                   // kuhn_poker.h:121
    info.information_state_tensor_size = 6 * n_ - 1;     // kuhn_poker.cc:395-401
    info.observation_tensor_size = 3 * n_ + 1;           // kuhn_poker.cc:403-410
    info.min_utility = -2; info.max_utility = (n_ - 1) * 2;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<KuhnState>(n_); }
 private:
  int n_;
};

}  // namespace
std::unique_ptr<Game> MakeKuhnPoker(const Params& p) { return std::make_unique<KuhnGame>(p); }
}  // namespace oracle
