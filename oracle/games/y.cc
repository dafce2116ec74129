// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/y/y.{h,cc}: an array of cells with the reference's union-find groups
// (parent / size / edge set per leader), independent of the flood fills of the CUDA rule core.
#include "../oracle.h"

namespace oracle {
namespace {

enum { kP1 = 0, kP2 = 1, kNone = 2, kInvalid = 3 };           // YPlayer, y.h:44-49

class YState : public State {
 public:
  explicit YState(int n) : n_(n), board_(n * n) {                                              // y.cc:112-124
    for (int i = 0; i < n * n; ++i) {
      const int x = i % n, y = i / n;
      const bool on = x + y < n;
      board_[i] = {on ? kNone : kInvalid, i, 1, on ? Edge(x, y) : 0};
    }
  }
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }       // y.h:124-126
  bool IsTerminal() const override { return outcome_ != kNone; }
  std::vector<int64_t> LegalActions() const override {                                         // y.cc:130-141
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < n_ * n_; ++c) if (board_[c].player == kNone) v.push_back(c);
    return v;
  }
  std::vector<double> Returns() const override {                                               // y.cc:214-218
    if (outcome_ == kP1) return {1.0, -1.0};
    if (outcome_ == kP2) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  std::string ToString() const override {                                                      // y.cc:147-212 (no colours)
    std::string s = " ";
    for (int x = 0; x < n_; ++x) { s += ' '; s += (char)('a' + x); }
    s += '\n';
    for (int y = 0; y < n_; ++y) {
      s += std::string(y + ((y + 1) < 10), ' ');
      s += std::to_string(y + 1);
      bool found_last = false;
      for (int x = 0; x < n_ - y; ++x) {
        const int xy = x + y * n_;
        if (found_last) { s += ']'; found_last = false; }
        else if (last_ == xy) { s += '['; found_last = true; }
        else s += ' ';
        const int p = board_[xy].player;
        if (p == kNone) s += '.';
        if (p == kP1) s += 'O';
        if (p == kP2) s += '@';
      }
      if (found_last) s += ']';
      s += '\n';
    }
    return s;
  }
  void ObservationTensor(int player, float* out) const override {                              // y.cc:232-258
    const int cells = n_ * n_;
    for (int i = 0; i < 3 * cells; ++i) out[i] = 0.f;
    for (int i = 0; i < cells; ++i) {
      const int p = board_[i].player;
      if (p == kInvalid) continue;
      const int plane = p == kNone ? 2 : (p == player ? 0 : 1);
      out[plane * cells + i] = 1.f;
    }
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<YState>(*this); }

 protected:
  void DoApplyAction(int64_t a) override {                                                     // y.cc:280-300
    if (a < 0 || a >= n_ * n_ || board_[a].player != kNone || outcome_ != kNone) { Fail("y: not an empty cell"); return; }
    const int x = (int)a % n_, y = (int)a / n_;
    last_ = (int)a;
    board_[a].player = cur_;
    static const int dx[6] = {0, 1, 1, 0, -1, -1}, dy[6] = {-1, -1, 0, 1, 1, 0};               // y.cc:60-64
    for (int d = 0; d < 6; ++d) {
      const int nx = x + dx[d], ny = y + dy[d];
      if (nx < 0 || ny < 0 || nx >= n_ || ny >= n_ || nx + ny >= n_) continue;
      if (board_[nx + ny * n_].player == cur_) Join((int)a, nx + ny * n_);
    }
    if (board_[Leader((int)a)].edge == 7) outcome_ = cur_;
    cur_ = 1 - cur_;
  }

 private:
  struct Cell { int player, parent, size, edge; };
  int Edge(int x, int y) const { return (x == 0 ? 1 : 0) | (y == 0 ? 2 : 0) | (x + y == n_ - 1 ? 4 : 0); }   // y.cc:103-108
  int Leader(int c) {                                                                          // y.cc:302-313
    while (board_[c].parent != c) c = board_[c].parent;
    return c;
  }
  void Join(int a, int b) {                                                                    // y.cc:315-335
    int la = Leader(a), lb = Leader(b);
    if (la == lb) return;
    if (board_[la].size < board_[lb].size) std::swap(la, lb);
    board_[lb].parent = la;
    board_[la].size += board_[lb].size;
    board_[la].edge |= board_[lb].edge;
  }
  int n_;
  std::vector<Cell> board_;
  int cur_ = kP1, outcome_ = kNone, last_ = -1;
};

class YGame : public Game {
 public:
  explicit YGame(const Params& p) {
    n_ = (int)p.get("board_size", 19);                  // y.h:39
    info.name = "y";
    info.num_distinct_actions = n_ * n_;
    info.max_game_length = n_ * (n_ + 1) / 2;
    info.observation_tensor_size = 3 * n_ * n_;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<YState>(n_); }
 private:
  int n_;
};

}  // namespace
std::unique_ptr<Game> MakeY(const Params& p) { return std::make_unique<YGame>(p); }
}  // namespace oracle
