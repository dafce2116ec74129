// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/othello/othello.{h,cc}: an 8x8 array of cells, moves found by walking the
// eight rays from a cell (independent of the bitboard fills of the CUDA rule core).
#include "../oracle.h"

namespace oracle {
namespace {

constexpr int kN = 8, kCells = 64, kPass = 64;              // othello.h:37-42
enum Cell { kEmpty = 0, kBlack = 1, kWhite = 2 };           // othello.h:45-49
constexpr int kDr[8] = {-1, 1, 0, 0, -1, -1, 1, 1};         // kUp, kDown, kLeft, kRight, kUpLeft, kUpRight, kDownLeft, kDownRight
constexpr int kDc[8] = {0, 0, -1, 1, -1, 1, -1, 1};         // (othello.cc:98-120)

class OthelloState : public State {
 public:
  OthelloState() {                                                                             // othello.cc:237-243
    for (int& c : board_) c = kEmpty;
    board_[3 * kN + 3] = kWhite; board_[3 * kN + 4] = kBlack;
    board_[4 * kN + 3] = kBlack; board_[4 * kN + 4] = kWhite;
  }
  int CurrentPlayer() const override { return cur_; }
  bool IsTerminal() const override { return cur_ == kTerminalPlayerId; }                      // othello.cc:272-274
  std::vector<int64_t> LegalActions() const override {                                         // othello.cc:219-224
    if (IsTerminal()) return {};
    std::vector<int64_t> v = Regular(cur_);
    if (v.empty()) v.push_back(kPass);
    return v;
  }
  std::vector<double> Returns() const override {                                               // othello.cc:276-284
    if (outcome_ == 0) return {1.0, -1.0};
    if (outcome_ == 1) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  std::string ToString() const override {                                                      // othello.cc:245-260
    const std::string cols = "  a b c d e f g h  ";
    std::string s = IsTerminal() ? std::string("Terminal State:\n") : std::string(cur_ == 0 ? "Black (x)" : "White (o)") + " to play:\n";
    s += cols + "\n";
    for (int r = 0; r < kN; ++r) {
      s += std::to_string(r + 1) + " ";
      for (int c = 0; c < kN; ++c) { s += "-xo"[board_[r * kN + c]]; s += ' '; }
      s += std::to_string(r + 1) + "\n";
    }
    return s + cols;
  }
  void ObservationTensor(int player, float* out) const override {                              // othello.cc:298-316
    for (int i = 0; i < 3 * kCells; ++i) out[i] = 0.f;
    const int mine = player == 0 ? kBlack : kWhite;
    for (int c = 0; c < kCells; ++c) out[(board_[c] == kEmpty ? 0 : board_[c] == mine ? 1 : 2) * kCells + c] = 1.f;
  }
  std::string InformationStateString(int) const override {                                     // HistoryString()
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<OthelloState>(*this); }

 protected:
  void DoApplyAction(int64_t a) override {                                                     // othello.cc:177-207
    if (IsTerminal()) { Fail("othello: move after the end of the game"); return; }
    if (a == kPass) {
      // the reference flips the player unconditionally; the batched API accepts a pass only where LegalActions() offers it
      if (!Regular(cur_).empty()) { Fail("othello: pass with a capture available"); return; }
      cur_ = 1 - cur_;
      return;
    }
    if (a < 0 || a >= kCells || !Valid(cur_, (int)a)) { Fail("othello: not a capturing move"); return; }
    const int mine = cur_ == 0 ? kBlack : kWhite;
    board_[a] = mine;
    for (int d = 0; d < 8; ++d) {
      const int steps = CountSteps(cur_, (int)a, d);
      int r = (int)a / kN + kDr[d], c = (int)a % kN + kDc[d];
      for (int i = 0; i < steps; ++i, r += kDr[d], c += kDc[d]) board_[r * kN + c] = mine;     // Capture(), :150-163
    }
    if (Regular(0).empty() && Regular(1).empty()) {                                            // NoValidActions(), :169-172
      int n0 = 0, n1 = 0;
      for (int c : board_) { n0 += c == kBlack; n1 += c == kWhite; }
      outcome_ = n0 > n1 ? 0 : n0 < n1 ? 1 : kInvalidPlayer;
      cur_ = kTerminalPlayerId;
    } else {
      cur_ = 1 - cur_;
    }
  }

 private:
  // opposing discs between `action` and the next disc of `player` along direction d; 0 if the ray ends first (:124-136)
  int CountSteps(int player, int action, int d) const {
    const int mine = player == 0 ? kBlack : kWhite;
    int r = action / kN + kDr[d], c = action % kN + kDc[d], count = 0;
    while (r >= 0 && r < kN && c >= 0 && c < kN) {
      if (board_[r * kN + c] == mine) return count;
      if (board_[r * kN + c] == kEmpty) return 0;
      ++count; r += kDr[d]; c += kDc[d];
    }
    return 0;
  }
  bool Valid(int player, int move) const {                                                     // :138-148, :174-176
    if (board_[move] != kEmpty) return false;
    for (int d = 0; d < 8; ++d) if (CountSteps(player, move, d) != 0) return true;
    return false;
  }
  std::vector<int64_t> Regular(int player) const {                                             // :209-217
    std::vector<int64_t> v;
    for (int c = 0; c < kCells; ++c) if (Valid(player, c)) v.push_back(c);
    return v;
  }
  int board_[kCells];
  int cur_ = 0, outcome_ = kInvalidPlayer;
};

class OthelloGame : public Game {
 public:
  OthelloGame() {
    info.name = "othello";
    info.num_distinct_actions = kCells + 1;      // othello.h:145
    info.max_game_length = 2 * kCells;           // othello.h:158
    info.observation_tensor_size = 3 * kCells;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<OthelloState>(); }
};

}  // namespace
std::unique_ptr<Game> MakeOthello(const Params&) { return std::make_unique<OthelloGame>(); }
}  // namespace oracle
