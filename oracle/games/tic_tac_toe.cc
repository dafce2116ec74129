// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/tic_tac_toe/tic_tac_toe.{h,cc}.
#include "../oracle.h"

namespace oracle {
namespace {

// tic_tac_toe.h:51-55 — cell enum order fixes the observation plane order.
enum Cell { kEmpty = 0, kNought = 1, kCross = 2 };

class TttState : public State {
 public:
  TttState() { for (int& c : board_) c = kEmpty; }

  // tic_tac_toe.h:104-106
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }

  // tic_tac_toe.cc:138-148
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < 9; ++c) if (board_[c] == kEmpty) v.push_back(c);
    return v;
  }

  // tic_tac_toe.cc:215-217
  bool IsTerminal() const override { return outcome_ != kInvalidPlayer || moves_ == 9; }

  // tic_tac_toe.cc:219-227
  std::vector<double> Returns() const override {
    if (HasLine(0)) return {1.0, -1.0};
    if (HasLine(1)) return {-1.0, 1.0};
    return {0.0, 0.0};
  }

  // tic_tac_toe.cc:165-176
  std::string ToString() const override {
    std::string s;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) s += ".ox"[board_[r * 3 + c]];
      if (r < 2) s += "\n";
    }
    return s;
  }

  // tic_tac_toe.cc:241-251 — TensorView<2>{3,9}, plane = cell enum.
  void ObservationTensor(int, float* out) const override {
    for (int i = 0; i < 27; ++i) out[i] = 0.f;
    for (int c = 0; c < 9; ++c) out[board_[c] * 9 + c] = 1.f;
  }
  std::string InformationStateString(int) const override {
    std::string s;   // HistoryString(): actions joined by ", " (spiel.h:560-562)
    for (size_t i = 0; i < history_.size(); ++i) {
      if (i) s += ", ";
      s += std::to_string(history_[i].second);
    }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<TttState>(*this); }

 protected:
  // tic_tac_toe.cc:128-136
  void DoApplyAction(int64_t a) override {
    if (a < 0 || a >= 9 || board_[a] != kEmpty) { Fail("ttt: cell not empty"); return; }
    board_[a] = cur_ == 0 ? kCross : kNought;      // PlayerToState, tic_tac_toe.cc:62-72
    if (HasLine(cur_)) outcome_ = cur_;
    cur_ = 1 - cur_;
    moves_ += 1;
  }

 private:
  // tic_tac_toe.cc:108-120 — the eight lines, spelled as index triples.
  bool HasLine(int player) const {
    static const int L[8][3] = {{0,1,2},{3,4,5},{6,7,8},{0,3,6},{1,4,7},{2,5,8},{0,4,8},{2,4,6}};
    int c = player == 0 ? kCross : kNought;
    for (auto& l : L) if (board_[l[0]] == c && board_[l[1]] == c && board_[l[2]] == c) return true;
    return false;
  }
  int board_[9];
  int cur_ = 0;
  int outcome_ = kInvalidPlayer;
  int moves_ = 0;
};

class TttGame : public Game {
 public:
  TttGame() {
    info.name = "tic_tac_toe";
    info.num_distinct_actions = 9;   // tic_tac_toe.h:136
    info.max_game_length = 9;        // tic_tac_toe.h:157
    info.observation_tensor_size = 27;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<TttState>(); }
};

}  // namespace
std::unique_ptr<Game> MakeTicTacToe(const Params&) { return std::make_unique<TttGame>(); }
}  // namespace oracle
