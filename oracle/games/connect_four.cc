// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/connect_four/connect_four.{h,cc}.
// Keeps the reference's cell-array board and its "scan every stone in four directions" line test,
// so it is independent of the bitboard shift-and test used by the CUDA kernels.
#include "../oracle.h"

namespace oracle {
namespace {

enum Cell { kEmpty = 0, kNought = 1, kCross = 2 };       // connect_four.h:52-56
enum Outcome { kP1 = 0, kP2 = 1, kUnknown = 2, kDraw = 3 };  // connect_four.h:58-63

struct C4Cfg { int rows, cols, x_in_row; bool ego; };

class C4State : public State {
 public:
  explicit C4State(const C4Cfg& g) : g_(g), board_(g.rows * g.cols, kEmpty) {}  // connect_four.cc:206-210

  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }  // :122-128

  // connect_four.cc:147-156
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < g_.cols; ++c) if (At(g_.rows - 1, c) == kEmpty) v.push_back(c);
    return v;
  }
  bool IsTerminal() const override { return outcome_ != kUnknown; }   // :277-279
  std::vector<double> Returns() const override {                      // :281-285
    if (outcome_ == kP1) return {1.0, -1.0};
    if (outcome_ == kP2) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  // connect_four.cc:212-222 — printed top row first.
  std::string ToString() const override {
    std::string s;
    for (int r = g_.rows - 1; r >= 0; --r) {
      for (int c = 0; c < g_.cols; ++c) s += ".ox"[At(r, c)];
      s += "\n";
    }
    return s;
  }
  // connect_four.cc:312-328 (+ PlayerRelative :299-310, StateToPlayer :75-86)
  void ObservationTensor(int player, float* out) const override {
    int n = g_.rows * g_.cols;
    for (int i = 0; i < 3 * n; ++i) out[i] = 0.f;
    for (int r = 0; r < g_.rows; ++r)
      for (int c = 0; c < g_.cols; ++c) {
        int cell = At(r, c), plane;
        if (g_.ego) {
          if (cell == kNought) plane = player == 0 ? 0 : 1;
          else if (cell == kCross) plane = player == 1 ? 0 : 1;
          else plane = 2;
        } else {
          plane = cell == kCross ? 0 : cell == kNought ? 1 : 2;
        }
        out[(plane * g_.rows + r) * g_.cols + c] = 1.f;
      }
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<C4State>(*this); }

 protected:
  // connect_four.cc:130-145
  void DoApplyAction(int64_t move) override {
    if (move < 0 || move >= g_.cols || At(g_.rows - 1, (int)move) != kEmpty) { Fail("c4: column full"); return; }
    int row = 0;
    while (At(row, (int)move) != kEmpty) ++row;
    board_[row * g_.cols + move] = cur_ == 0 ? kCross : kNought;
    if (HasLine(cur_)) outcome_ = cur_;
    else if (IsFull()) outcome_ = kDraw;
    cur_ = 1 - cur_;
  }

 private:
  int At(int r, int c) const { return board_[r * g_.cols + c]; }
  // connect_four.cc:170-185
  bool LineFromDir(int player, int row, int col, int dr, int dc) const {
    int k = g_.x_in_row;
    if (row + (k - 1) * dr >= g_.rows || col + (k - 1) * dc >= g_.cols ||
        row + (k - 1) * dr < 0 || col + (k - 1) * dc < 0) return false;
    int want = player == 0 ? kCross : kNought;
    for (int i = 0; i < k; ++i) {
      if (At(row, col) != want) return false;
      row += dr; col += dc;
    }
    return true;
  }
  // connect_four.cc:163-168, 187-196
  bool HasLine(int player) const {
    int want = player == 0 ? kCross : kNought;
    for (int c = 0; c < g_.cols; ++c)
      for (int r = 0; r < g_.rows; ++r)
        if (At(r, c) == want &&
            (LineFromDir(player, r, c, 0, 1) || LineFromDir(player, r, c, -1, -1) ||
             LineFromDir(player, r, c, -1, 0) || LineFromDir(player, r, c, -1, 1)))
          return true;
    return false;
  }
  bool IsFull() const {   // :198-204
    for (int c = 0; c < g_.cols; ++c) if (At(g_.rows - 1, c) == kEmpty) return false;
    return true;
  }
  C4Cfg g_;
  std::vector<int> board_;
  int cur_ = 0;
  int outcome_ = kUnknown;
};

class C4Game : public Game {
 public:
  explicit C4Game(const Params& p) {
    // defaults: connect_four.h:45-50
    cfg_.rows = (int)p.get("rows", 6);
    cfg_.cols = (int)p.get("columns", 7);
    cfg_.x_in_row = (int)p.get("x_in_row", 4);
    cfg_.ego = p.get("egocentric_obs_tensor", 0) != 0;
    info.name = "connect_four";
    info.num_distinct_actions = cfg_.cols;              // connect_four.h:179
    info.max_game_length = cfg_.rows * cfg_.cols;       // connect_four.h:200
    info.observation_tensor_size = 3 * cfg_.rows * cfg_.cols;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<C4State>(cfg_); }
 private:
  C4Cfg cfg_;
};

}  // namespace
std::unique_ptr<Game> MakeConnectFour(const Params& p) { return std::make_unique<C4Game>(p); }
}  // namespace oracle
