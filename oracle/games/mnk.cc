// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/mnk/mnk.{h,cc} (m columns, n rows, k in a row; default 15, 15, 5).
#include "../oracle.h"

namespace oracle {
namespace {

enum Cell { kEmpty = 0, kNought = 1, kCross = 2 };     // mnk.h:38-42: fixes the observation plane order

class MnkState : public State {
 public:
  MnkState(int rows, int cols, int k) : rows_(rows), cols_(cols), k_(k), board_(rows * cols, kEmpty) {}
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }        // mnk.h:51-53
  std::vector<int64_t> LegalActions() const override {                                         // mnk.cc:148-160
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < rows_ * cols_; ++c) if (board_[c] == kEmpty) v.push_back(c);
    return v;
  }
  bool IsTerminal() const override { return outcome_ != kInvalidPlayer || moves_ == rows_ * cols_; }   // mnk.cc:207-209
  std::vector<double> Returns() const override {                                               // mnk.cc:211-219
    if (HasLine(0)) return {1.0, -1.0};
    if (HasLine(1)) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  std::string ToString() const override {                                                      // mnk.cc:193-205
    std::string s;
    for (int r = 0; r < rows_; ++r) {
      for (int c = 0; c < cols_; ++c) s += ".ox"[board_[r * cols_ + c]];
      if (r < rows_ - 1) s += "\n";
    }
    return s;
  }
  void ObservationTensor(int, float* out) const override {                                     // mnk.cc:233-246
    const int cells = rows_ * cols_;
    for (int i = 0; i < 3 * cells; ++i) out[i] = 0.f;
    for (int c = 0; c < cells; ++c) out[board_[c] * cells + c] = 1.f;
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<MnkState>(*this); }

 protected:
  void DoApplyAction(int64_t a) override {                                                     // mnk.cc:117-126
    if (a < 0 || a >= rows_ * cols_ || board_[a] != kEmpty) { Fail("mnk: cell not empty"); return; }
    board_[a] = cur_ == 0 ? kCross : kNought;
    if (HasLine(cur_)) outcome_ = cur_;
    cur_ = 1 - cur_;
    moves_ += 1;
  }

 private:
  // BoardHasLine, mnk.cc:92-115: k consecutive cells of the player's colour from any cell in any of the eight directions
  bool HasLine(int player) const {
    const int want = player == 0 ? kCross : kNought;
    for (int r = 0; r < rows_; ++r)
      for (int c = 0; c < cols_; ++c)
        for (int dr = -1; dr <= 1; ++dr)
          for (int dc = -1; dc <= 1; ++dc) {
            if (!dr && !dc) continue;
            int count = 0, rr = r, cc = c;
            for (int i = 0; i < k_ && rr >= 0 && rr < rows_ && cc >= 0 && cc < cols_; ++i, rr += dr, cc += dc)
              count += board_[rr * cols_ + cc] == want;
            if (count == k_) return true;
          }
    return false;
  }
  int rows_, cols_, k_;
  std::vector<int> board_;
  int cur_ = 0, outcome_ = kInvalidPlayer, moves_ = 0;
};

class MnkGame : public Game {
 public:
  explicit MnkGame(const Params& p) {
    cols_ = (int)p.get("m", 15); rows_ = (int)p.get("n", 15); k_ = (int)p.get("k", 5);       // mnk.h:34-36
    info.name = "mnk";
    info.num_distinct_actions = rows_ * cols_;
    info.max_game_length = rows_ * cols_;
    info.observation_tensor_size = 3 * rows_ * cols_;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<MnkState>(rows_, cols_, k_); }
 private:
  int rows_, cols_, k_;
};

}  // namespace
std::unique_ptr<Game> MakeMnk(const Params& p) { return std::make_unique<MnkGame>(p); }
}  // namespace oracle
