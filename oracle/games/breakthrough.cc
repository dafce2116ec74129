// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/breakthrough/breakthrough.{h,cc}.
#include "../oracle.h"

namespace oracle {
namespace {

enum Cell { kEmpty = 0, kBlack = 1, kWhite = 2 };
// breakthrough.cc:36-40 — black (player 0) moves row+1, white (player 1) row-1.
const int kDR[6] = {1, 1, 1, -1, -1, -1};
const int kDC[6] = {-1, 0, 1, -1, 0, 1};

class BtState : public State {
 public:
  // breakthrough.cc:121-144
  BtState(int rows, int cols) : rows_(rows), cols_(cols), board_(rows * cols, kEmpty) {
    for (int r = 0; r < rows_; ++r)
      for (int c = 0; c < cols_; ++c) {
        if (r == 0 || (rows_ >= 6 && r == 1)) board_[r * cols_ + c] = kBlack;
        else if (r == rows_ - 1 || (rows_ >= 6 && r == rows_ - 2)) board_[r * cols_ + c] = kWhite;
      }
    pieces_[0] = pieces_[1] = cols_ * (rows_ >= 6 ? 2 : 1);
  }
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }

  // breakthrough.cc:219-258; action rank = ((r*cols+c)*6+dir)*2+capture (spiel_utils.cc:50-63)
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    int mine = cur_ == 0 ? kBlack : kWhite, theirs = cur_ == 0 ? kWhite : kBlack;
    for (int r = 0; r < rows_; ++r)
      for (int c = 0; c < cols_; ++c) {
        if (At(r, c) != mine) continue;
        for (int o = 0; o < 3; ++o) {
          int dir = cur_ * 3 + o, rp = r + kDR[dir], cp = c + kDC[dir];
          if (!In(rp, cp)) continue;
          int64_t base = ((int64_t)(r * cols_ + c) * 6 + dir) * 2;
          if (At(rp, cp) == kEmpty) v.push_back(base);
          else if ((o == 0 || o == 2) && At(rp, cp) == theirs) v.push_back(base + 1);
        }
      }
    return v;
  }
  // Playout candidates (see oracle.h): for every own piece in ascending cell order, its three forward moves
  // (left diagonal, straight, right diagonal); a move that leaves the board is listed as -1 (never legal), a
  // diagonal onto an enemy piece as the capture id, anything else as the plain-move id (legal iff the target is empty).
  std::vector<int64_t> RolloutCandidates() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    int mine = cur_ == 0 ? kBlack : kWhite, theirs = cur_ == 0 ? kWhite : kBlack;
    for (int r = 0; r < rows_; ++r)
      for (int c = 0; c < cols_; ++c) {
        if (At(r, c) != mine) continue;
        for (int o = 0; o < 3; ++o) {
          int dir = cur_ * 3 + o, rp = r + kDR[dir], cp = c + kDC[dir];
          if (!In(rp, cp)) { v.push_back(-1); continue; }
          int64_t base = ((int64_t)(r * cols_ + c) * 6 + dir) * 2;
          v.push_back(base + ((o != 1 && At(rp, cp) == theirs) ? 1 : 0));
        }
      }
    return v;
  }
  bool IsTerminal() const override { return winner_ >= 0 || pieces_[0] == 0 || pieces_[1] == 0; }  // :308-310
  std::vector<double> Returns() const override {   // :312-320
    if (winner_ == 0 || pieces_[1] == 0) return {1.0, -1.0};
    if (winner_ == 1 || pieces_[0] == 0) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  // breakthrough.cc:264-284
  std::string ToString() const override {
    std::string s;
    for (int r = 0; r < rows_; ++r) {
      s += (char)('1' + (rows_ - 1 - r));
      for (int c = 0; c < cols_; ++c) s += ".bw"[At(r, c)];
      s += "\n";
    }
    s += " ";
    for (int c = 0; c < cols_; ++c) s += (char)('a' + c);
    s += "\n";
    return s;
  }
  // breakthrough.cc:286-306, 328-342 — planes 0=black, 1=white, 2=empty
  void ObservationTensor(int, float* out) const override {
    int n = rows_ * cols_;
    for (int i = 0; i < 3 * n; ++i) out[i] = 0.f;
    for (int i = 0; i < n; ++i) {
      int plane = board_[i] == kBlack ? 0 : board_[i] == kWhite ? 1 : 2;
      out[plane * n + i] = 1.f;
    }
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<BtState>(*this); }

 protected:
  // breakthrough.cc:154-194
  void DoApplyAction(int64_t a) override {
    if (a < 0 || a >= (int64_t)rows_ * cols_ * 12) { Fail("bt: action out of range"); return; }
    int cap = a % 2, dir = (a / 2) % 6, c1 = (a / 12) % cols_, r1 = (int)(a / 12 / cols_);
    int r2 = r1 + kDR[dir], c2 = c1 + kDC[dir];
    if (!In(r1, c1) || !In(r2, c2)) { Fail("bt: out of bounds"); return; }
    if (At(r2, c2) == kWhite) {
      pieces_[1]--;
      if (At(r1, c1) != kBlack || cur_ != 0) { Fail("bt: bad capture"); return; }
    } else if (At(r2, c2) == kBlack) {
      pieces_[0]--;
      if (At(r1, c1) != kWhite || cur_ != 1) { Fail("bt: bad capture"); return; }
    }
    if (cap) {
      int from = At(r1, c1);
      int opp = from == kBlack ? kWhite : from == kWhite ? kBlack : -1;
      if (At(r2, c2) != opp) { Fail("bt: capture flag without capture"); return; }
    }
    board_[r2 * cols_ + c2] = At(r1, c1);
    board_[r1 * cols_ + c1] = kEmpty;
    if (cur_ == 0 && r2 == rows_ - 1) winner_ = 0;
    else if (cur_ == 1 && r2 == 0) winner_ = 1;
    cur_ = 1 - cur_;
    total_moves_++;
  }

 private:
  int At(int r, int c) const { return board_[r * cols_ + c]; }
  bool In(int r, int c) const { return r >= 0 && r < rows_ && c >= 0 && c < cols_; }
  int rows_, cols_;
  std::vector<int> board_;
  int pieces_[2];
  int winner_ = kInvalidPlayer;
  int cur_ = 0;
  int total_moves_ = 0;
};

class BtGame : public Game {
 public:
  explicit BtGame(const Params& p) {
    rows_ = (int)p.get("rows", 8);
    cols_ = (int)p.get("columns", 8);
    info.name = "breakthrough";
    info.num_distinct_actions = rows_ * cols_ * 6 * 2;             // breakthrough.cc:388-390
    info.max_game_length = 2 * (2 * rows_ - 3) * cols_ + 1;        // breakthrough.h:118-120
    info.observation_tensor_size = 3 * rows_ * cols_;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<BtState>(rows_, cols_); }
 private:
  int rows_, cols_;
};

}  // namespace
std::unique_ptr<Game> MakeBreakthrough(const Params& p) { return std::make_unique<BtGame>(p); }
}  // namespace oracle
