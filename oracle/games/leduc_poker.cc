// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/leduc_poker/leduc_poker.{h,cc}
// (players, starting_player parameters; action_mapping / suit_isomorphism are not restated).
#include <cstdio>

#include "../oracle.h"

namespace oracle {
namespace {

constexpr int kInvalidCard = -10000;     // leduc_poker.h:60
constexpr int kStartingMoney = 100;      // leduc_poker.h:68
enum { kFold = 0, kCall = 1, kRaise = 2 };

std::string Num(double v) {              // absl::StrCat(double) prints like "%g"
  char b[64];
  snprintf(b, sizeof b, "%g", v);
  return b;
}
template <typename V>
std::string Join(const V& v, const char* sep) {
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) { if (i) s += sep; s += Num((double)v[i]); }
  return s;
}

class LeducState : public State {
 public:
  // leduc_poker.cc:239-282
  LeducState(int n, int starting_player)
      : n_(n), starting_player_(starting_player), pot_(n), deck_size_((n + 1) * 2), remaining_(n),
        winner_(n, false), priv_(n, kInvalidCard), money_(n, kStartingMoney - 1), ante_(n, 1), folded_(n, false) {
    deck_.resize(deck_size_);
    for (int i = 0; i < deck_size_; ++i) deck_[i] = i;
  }
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }

  // leduc_poker.cc:416-457
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    if (IsChanceNode()) {
      for (int c = 0; c < (int)deck_.size(); ++c) if (deck_[c] != kInvalidCard) v.push_back(c);
      return v;
    }
    if (stakes_ > ante_[cur_]) v.push_back(kFold);
    v.push_back(kCall);
    if (num_raises_ < 2) v.push_back(kRaise);
    return v;
  }
  bool IsTerminal() const override { return remaining_ == 1 || (round_ == 2 && ReadyForNextRound()); }  // :498-500
  std::vector<double> Returns() const override {       // :502-514
    std::vector<double> r(n_, 0.0);
    if (!IsTerminal()) return r;
    for (int p = 0; p < n_; ++p) r[p] = money_[p] - kStartingMoney;
    return r;
  }
  // leduc_poker.cc:463-496
  std::string ToString() const override {
    static const char* kNames[3] = {"Fold", "Call", "Raise"};
    std::string s = "Round: " + std::to_string(round_) + "\nPlayer: " + std::to_string(cur_) +
                    "\nPot: " + std::to_string(pot_) + "\nMoney (player_0 player_1" + (n_ > 2 ? " [...]):" : "):");
    for (int p = 0; p < n_; ++p) s += " " + Num(money_[p]);
    s += std::string("\nCards (public player_0 player_1") + (n_ > 2 ? " [...]): " : "): ") +
         std::to_string(public_card_) + " ";
    for (int p = 0; p < n_; ++p) s += std::to_string(priv_[p]) + " ";
    s += "\nRound 1 sequence: ";
    for (size_t i = 0; i < r1_.size(); ++i) { if (i) s += ", "; s += kNames[r1_[i]]; }
    s += "\nRound 2 sequence: ";
    for (size_t i = 0; i < r2_.size(); ++i) { if (i) s += ", "; s += kNames[r2_[i]]; }
    s += "\n";
    return s;
  }
  // LeducObserver::StringFrom, leduc_poker.cc:198-239
  std::string ObsString(int player, bool perfect_recall) const {
    std::string s = "[Observer: " + std::to_string(player) + "][Private: " + std::to_string(priv_[player]) + "]";
    s += "[Round " + std::to_string(round_) + "][Player: " + std::to_string(cur_) + "][Pot: " +
         std::to_string(pot_) + "][Money: " + Join(money_, " ") + "]";
    if (public_card_ != kInvalidCard) s += "[Public: " + std::to_string(public_card_) + "]";
    if (perfect_recall) s += "[Round1: " + Join(r1_, " ") + "][Round2: " + Join(r2_, " ") + "]";
    else s += "[Ante: " + Join(ante_, " ") + "]";
    return s;
  }
  std::string InformationStateString(int p) const override { return ObsString(p, true); }
  std::string ObservationString(int p) const override { return ObsString(p, false); }

  // LeducObserver::WriteTensor, leduc_poker.cc:92-192; field order = ContiguousAllocator Get() order.
  void InformationStateTensor(int player, float* out) const override {
    int cards = (int)deck_.size(), mb = 3 * n_ - 2;
    int sz = n_ + 2 * cards + 2 * mb * 2;
    for (int i = 0; i < sz; ++i) out[i] = 0.f;
    out[player] = 1.f;
    if (priv_[player] != kInvalidCard) out[n_ + priv_[player]] = 1.f;
    if (public_card_ != kInvalidCard) out[n_ + cards + public_card_] = 1.f;
    float* bet = out + n_ + 2 * cards;      // {2, max_bets_per_round, 2}; call = 10, raise = 01, fold = 00
    for (int round = 0; round < 2; ++round) {
      const auto& seq = round == 0 ? r1_ : r2_;
      for (size_t i = 0; i < seq.size(); ++i) {
        if (seq[i] == kCall) bet[(round * mb + i) * 2 + 0] = 1.f;
        else if (seq[i] == kRaise) bet[(round * mb + i) * 2 + 1] = 1.f;
      }
    }
  }
  void ObservationTensor(int player, float* out) const override {
    int cards = (int)deck_.size();
    int sz = n_ + 2 * cards + n_;
    for (int i = 0; i < sz; ++i) out[i] = 0.f;
    out[player] = 1.f;
    if (priv_[player] != kInvalidCard) out[n_ + priv_[player]] = 1.f;
    if (public_card_ != kInvalidCard) out[n_ + cards + public_card_] = 1.f;
    for (int p = 0; p < n_; ++p) out[n_ + 2 * cards + p] = (float)ante_[p];
  }
  // leduc_poker.cc:546-571
  std::vector<std::pair<int64_t, double>> ChanceOutcomes() const override {
    std::vector<std::pair<int64_t, double>> o;
    double p = 1.0 / deck_size_;
    for (int c = 0; c < (int)deck_.size(); ++c) if (deck_[c] != kInvalidCard) o.push_back({c, p});
    return o;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<LeducState>(*this); }

 protected:
  // leduc_poker.cc:298-414
  void DoApplyAction(int64_t move) override {
    if (IsChanceNode()) {
      if (move < 0 || move >= (int64_t)deck_.size() || deck_[move] == kInvalidCard) { Fail("leduc: bad card"); return; }
      if (dealt_ < n_) {
        priv_[dealt_] = deck_[move];          // SetPrivate, :722-744
        deck_[move] = kInvalidCard;
        --deck_size_;
        ++dealt_;
        if (dealt_ == n_) cur_ = starting_player_;
      } else {
        public_card_ = deck_[move];
        deck_[move] = kInvalidCard;
        --deck_size_;
        cur_ = NextPlayer();
      }
      return;
    }
    if (move == kFold) {
      Append(kFold);
      folded_[cur_] = true;
      remaining_--;
      if (IsTerminal()) ResolveWinner();
      else if (ReadyForNextRound()) NewRound();
      else cur_ = NextPlayer();
    } else if (move == kCall) {
      if (stakes_ < ante_[cur_]) { Fail("leduc: stakes < ante"); return; }
      AnteUp(cur_, stakes_ - ante_[cur_]);
      num_calls_++;
      Append(kCall);
      if (IsTerminal()) ResolveWinner();
      else if (ReadyForNextRound()) NewRound();
      else cur_ = NextPlayer();
    } else if (move == kRaise) {
      if (num_raises_ >= 2) { Fail("leduc: too many raises"); return; }
      int call_amount = stakes_ - ante_[cur_];
      if (call_amount > 0) AnteUp(cur_, call_amount);
      int raise = round_ == 1 ? 2 : 4;
      stakes_ += raise;
      AnteUp(cur_, raise);
      num_raises_++;
      num_calls_ = 0;
      Append(kRaise);
      if (IsTerminal()) ResolveWinner();
      else cur_ = NextPlayer();
    } else {
      Fail("leduc: invalid move");
    }
  }

 private:
  bool ReadyForNextRound() const {       // :680-683
    return (num_raises_ == 0 && num_calls_ == remaining_) || (num_raises_ > 0 && num_calls_ == remaining_ - 1);
  }
  void NewRound() { round_++; num_raises_ = 0; num_calls_ = 0; cur_ = kChancePlayerId; }   // :685-691
  void Append(int m) { (round_ == 1 ? r1_ : r2_).push_back(m); }
  void AnteUp(int p, int amount) { pot_ += amount; ante_[p] += amount; money_[p] -= amount; }  // :702-706
  int NextPlayer() const {               // :573-591
    int from = cur_ == kChancePlayerId ? (starting_player_ + n_ - 1) % n_ : cur_;
    for (int i = 1; i <= n_; ++i) { int p = (from + i) % n_; if (!folded_[p]) return p; }
    return -1;
  }
  int RankHand(int player) const {       // :593-626
    int lo = public_card_, hi = priv_[player];
    if (lo > hi) std::swap(lo, hi);
    int nc = (int)deck_.size();
    if (lo % 2 == 0 && hi == lo + 1) return nc * nc + lo;
    return (hi / 2) * nc + (lo / 2);
  }
  void ResolveWinner() {                 // :628-678
    if (remaining_ == 1) {
      for (int p = 0; p < n_; ++p)
        if (!folded_[p]) { winner_[p] = true; money_[p] += pot_; pot_ = 0; return; }
    } else {
      int best = -1, nw = 0;
      std::fill(winner_.begin(), winner_.end(), false);
      for (int p = 0; p < n_; ++p) {
        if (folded_[p]) continue;
        int rank = RankHand(p);
        if (rank > best) { best = rank; std::fill(winner_.begin(), winner_.end(), false); winner_[p] = true; nw = 1; }
        else if (rank == best) { winner_[p] = true; nw++; }
      }
      for (int p = 0; p < n_; ++p) if (winner_[p]) money_[p] += static_cast<double>(pot_) / nw;
      pot_ = 0;
    }
  }
  int n_, starting_player_;
  int cur_ = kChancePlayerId;
  int num_calls_ = 0, num_raises_ = 0, round_ = 1, stakes_ = 1;
  int pot_;
  int public_card_ = kInvalidCard;
  int deck_size_;
  int dealt_ = 0;
  int remaining_;
  std::vector<bool> winner_;
  std::vector<int> priv_;
  std::vector<double> money_;
  std::vector<int> ante_;
  std::vector<bool> folded_;
  std::vector<int> deck_;
  std::vector<int> r1_, r2_;
};

class LeducGame : public Game {
 public:
  explicit LeducGame(const Params& p) {
    n_ = (int)p.get("players", 2);
    sp_ = (int)p.get("starting_player", 0);
    int cards = (n_ + 1) * 2, mgl = 2 * (3 * n_ - 2);
    info.name = "leduc_poker";
    info.num_players = n_;
    info.num_distinct_actions = 3;
    info.max_chance_outcomes = cards;
    info.max_game_length = mgl;                                         // leduc_poker.h:233-241
    info.information_state_tensor_size = n_ + cards * 2 + mgl * 2;      // leduc_poker.cc:811-820
    info.observation_tensor_size = n_ + cards * 2 + n_;                 // leduc_poker.cc:822-831
    info.max_utility = (n_ - 1) * (2 * 2 + 2 * 4 + 1);                  // leduc_poker.cc:833-841
    info.min_utility = -(2 * 2 + 2 * 4 + 1);
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<LeducState>(n_, sp_); }
 private:
  int n_, sp_;
};

}  // namespace
std::unique_ptr<Game> MakeLeducPoker(const Params& p) { return std::make_unique<LeducGame>(p); }
}  // namespace oracle
