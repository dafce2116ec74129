// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/hex/hex.{h,cc}.
// Keeps the reference's per-cell label array, neighbour lists and stack flood fill.
#include "../oracle.h"

namespace oracle {
namespace {

// hex.h:68-78 — nine cell labels; value + 4 is the observation plane.
enum Cell {
  kEmpty = 0, kWhiteWest = -3, kWhiteEast = -2, kWhiteWin = -4, kWhite = -1,
  kBlackNorth = 3, kBlackSouth = 2, kBlackWin = 4, kBlack = 1,
};

struct HexCfg { int cols, rows; bool swap, plain_obs; };

class HexState : public State {
 public:
  explicit HexState(const HexCfg& g) : g_(g), board_(g.cols * g.rows, kEmpty) {}
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }

  // hex.cc:280-293
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < (int)board_.size(); ++c) if (board_[c] == kEmpty) v.push_back(c);
    if (g_.swap && history_.size() == 1 && cur_ == 1) v.push_back(g_.cols * g_.rows);
    return v;
  }
  bool IsTerminal() const override { return result_ != 0; }          // hex.cc:361
  std::vector<double> Returns() const override { return {result_, -result_}; }   // hex.cc:363-365 (−0.0!)

  // hex.cc:343-359 (standard string_rep)
  std::string ToString() const override {
    std::string s;
    int line = 0;
    for (int cell = 0; cell < (int)board_.size(); ++cell) {
      if (cell && cell % g_.cols == 0) { s += "\n"; ++line; s += std::string(line, ' '); }
      s += board_[cell] == kEmpty ? "." : board_[cell] < 0 ? "o" : "x";
      s += " ";
    }
    return s;
  }
  // hex.cc:379-398
  void ObservationTensor(int, float* out) const override {
    int n = (int)board_.size();
    if (g_.plain_obs) {
      for (int i = 0; i < 3 * n; ++i) out[i] = 0.f;
      for (int cell = 0; cell < n; ++cell) {
        int plane = board_[cell] == kEmpty ? 2 : board_[cell] < 0 ? 1 : 0;   // hex.cc:76-93
        // TensorView<3>{3, num_cols, num_rows} indexed {plane, cell / num_cols, cell % num_cols}
        out[(plane * g_.cols + cell / g_.cols) * g_.rows + cell % g_.cols] = 1.f;
      }
    } else {
      for (int i = 0; i < 9 * n; ++i) out[i] = 0.f;
      for (int cell = 0; cell < n; ++cell) out[(board_[cell] + 4) * n + cell] = 1.f;
    }
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<HexState>(*this); }

 protected:
  // hex.cc:229-278
  void DoApplyAction(int64_t move) override {
    if (g_.swap && move == g_.cols * g_.rows) {
      if (history_.size() != 1 || cur_ != 1) { Fail("hex: swap not allowed"); return; }
      int first = (int)history_[0].second;
      board_[first] = kEmpty;
      int r = first / g_.cols, c = first % g_.cols;
      int mirrored = c * g_.cols + r;
      // hex.cc:238 indexes the board with this value unchecked; on boards with more columns than rows it can lie outside
      // the board (undefined behaviour in the reference) — reported as an error here instead of corrupting memory.
      if (mirrored >= (int)board_.size()) { Fail("hex: swap mirrors the first stone outside a non-square board"); return; }
      board_[mirrored] = LabelFor(1, mirrored);
      cur_ = 0;
      return;
    }
    if (move < 0 || move >= (int64_t)board_.size() || board_[move] != kEmpty) { Fail("hex: cell not empty"); return; }
    int label = LabelFor(cur_, (int)move);
    board_[move] = label;
    if (label == kBlackWin) result_ = 1;
    else if (label == kWhiteWin) result_ = -1;
    else if (label != kBlack && label != kWhite) {
      int plain = cur_ == 0 ? kBlack : kWhite;
      std::vector<int> stack = {(int)move};
      while (!stack.empty()) {
        int cell = stack.back();
        stack.pop_back();
        for (int nb : Adjacent(cell))
          if (board_[nb] == plain) { board_[nb] = label; stack.push_back(nb); }
      }
    }
    cur_ = 1 - cur_;
  }

 private:
  // hex.cc:316-329 — N, NE, E, S, SW, W
  std::vector<int> Adjacent(int cell) const {
    std::vector<int> nb;
    int n = (int)board_.size();
    bool north = cell < g_.cols, south = cell >= n - g_.cols;
    bool west = cell % g_.cols == 0, east = cell % g_.cols == g_.cols - 1;
    if (!north) nb.push_back(cell - g_.cols);
    if (!north && !east) nb.push_back(cell - g_.cols + 1);
    if (!east) nb.push_back(cell + 1);
    if (!south) nb.push_back(cell + g_.cols);
    if (!south && !west) nb.push_back(cell + g_.cols - 1);
    if (!west) nb.push_back(cell - 1);
    return nb;
  }
  // hex.cc:108-171 — note the `else if` on the own-edge test (first row wins over last row).
  int LabelFor(int player, int move) const {
    int n = (int)board_.size();
    if (player == 0) {
      bool north = false, south = false;
      if (move < g_.cols) north = true;
      else if (move >= n - g_.cols) south = true;
      for (int nb : Adjacent(move)) {
        if (board_[nb] == kBlackNorth) north = true;
        else if (board_[nb] == kBlackSouth) south = true;
      }
      return north && south ? kBlackWin : north ? kBlackNorth : south ? kBlackSouth : kBlack;
    }
    bool west = false, east = false;
    if (move % g_.cols == 0) west = true;
    else if (move % g_.cols == g_.cols - 1) east = true;
    for (int nb : Adjacent(move)) {
      if (board_[nb] == kWhiteWest) west = true;
      else if (board_[nb] == kWhiteEast) east = true;
    }
    return west && east ? kWhiteWin : west ? kWhiteWest : east ? kWhiteEast : kWhite;
  }
  HexCfg g_;
  std::vector<int> board_;
  int cur_ = 0;
  double result_ = 0;
};

class HexGame : public Game {
 public:
  explicit HexGame(const Params& p) {
    int bs = (int)p.get("board_size", 11);     // hex.h:41-44 defaults
    cfg_.cols = (int)p.get("num_cols", bs);
    cfg_.rows = (int)p.get("num_rows", bs);
    cfg_.swap = p.get("swap", 0) != 0;
    cfg_.plain_obs = p.get("plain_obs_tensor", 0) != 0;
    info.name = "hex";
    info.num_distinct_actions = cfg_.cols * cfg_.rows + (cfg_.swap ? 1 : 0);   // hex.h:135-137
    info.max_game_length = cfg_.cols * cfg_.rows;                              // hex.h:147
    info.observation_tensor_size = (cfg_.plain_obs ? 3 : 9) * cfg_.cols * cfg_.rows;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<HexState>(cfg_); }
 private:
  HexCfg cfg_;
};

}  // namespace
std::unique_ptr<Game> MakeHex(const Params& p) { return std::make_unique<HexGame>(p); }
}  // namespace oracle
