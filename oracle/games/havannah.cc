// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/havannah/havannah.{h,cc}: an array of cells with the reference's union-find
// groups (parent / size / corner set / edge set per leader), its neighbour loop with the skip rule, and the recursive ring
// search — independent of the flood fills and the explicit-stack search of the CUDA rule core.
#include "../oracle.h"

namespace oracle {
namespace {

enum { kP1 = 0, kP2 = 1, kNone = 2, kDraw = 3, kInvalid = 4 };          // HavannahPlayer, havannah.h:44-50
constexpr int kDx[6] = {-1, 0, 1, 1, 0, -1}, kDy[6] = {-1, -1, 0, 1, 1, 0};   // havannah.cc:74-78

class HavannahState : public State {
 public:
  HavannahState(int size, bool swap) : size_(size), d_(2 * size - 1), swap_(swap), board_(d_ * d_) {   // havannah.cc:163-180
    valid_ = d_ * d_ - size * (size - 1);
    for (int i = 0; i < d_ * d_; ++i) {
      const int x = i % d_, y = i / d_;
      const bool on = OnBoard(x, y);
      board_[i] = {on ? kNone : kInvalid, false, i, 1, on ? Corner(x, y) : 0, on ? Edge(x, y) : 0};
    }
  }
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : cur_; }
  bool IsTerminal() const override { return outcome_ != kNone; }
  std::vector<int64_t> LegalActions() const override {                                         // havannah.cc:186-201
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int c = 0; c < d_ * d_; ++c) if (board_[c].player == kNone || (AllowSwap() && c == last_)) v.push_back(c);
    return v;
  }
  std::vector<double> Returns() const override {                                               // havannah.cc:281-286
    if (outcome_ == kP1) return {1.0, -1.0};
    if (outcome_ == kP2) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  std::string ToString() const override {                                                      // havannah.cc:212-279 (no colours)
    std::string s(size_ + 3, ' ');
    for (int x = 0; x < size_; ++x) { s += ' '; s += (char)('a' + x); }
    s += '\n';
    for (int y = 0; y < d_; ++y) {
      s += std::string(std::abs(size_ - 1 - y) + 1 + ((y + 1) < 10), ' ');
      s += std::to_string(y + 1);
      bool found_last = false;
      const int start_x = y < size_ ? 0 : y - size_ + 1, end_x = y < size_ ? size_ + y : d_;
      for (int x = start_x; x < end_x; ++x) {
        const int xy = x + y * d_;
        if (found_last) { s += ']'; found_last = false; }
        else if (last_ == xy) { s += '['; found_last = true; }
        else s += ' ';
        const int p = board_[xy].player;
        if (p == kNone) s += '.';
        if (p == kP1) s += 'O';
        if (p == kP2) s += '@';
      }
      if (found_last) s += ']';
      if (y < size_ - 1) { s += ' '; s += (char)('a' + size_ + y); }
      s += '\n';
    }
    return s;
  }
  void ObservationTensor(int player, float* out) const override {                              // havannah.cc:296-322
    const int cells = d_ * d_;
    for (int i = 0; i < 3 * cells; ++i) out[i] = 0.f;
    for (int i = 0; i < cells; ++i) {
      const int p = board_[i].player;
      if (p >= 3) continue;
      out[(p == kNone ? 2 : (p == player ? 0 : 1)) * cells + i] = 1.f;
    }
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<HavannahState>(*this); }

 protected:
  void DoApplyAction(int64_t a) override {                                                     // havannah.cc:324-359
    if (outcome_ != kNone || a < 0 || a >= d_ * d_ || board_[a].player == kInvalid) { Fail("havannah: not a cell"); return; }
    if (last_ == (int)a) {
      if (!AllowSwap()) { Fail("havannah: occupied"); return; }
    } else {
      if (board_[a].player != kNone) { Fail("havannah: occupied"); return; }
      ++moves_;
      last_ = (int)a;
    }
    board_[a].player = cur_;
    const int x = (int)a % d_, y = (int)a / d_;
    bool alreadyjoined = false, skip = false;
    for (int dir = 0; dir < 6; ++dir) {
      if (skip) { skip = false; continue; }
      const int nx = x + kDx[dir], ny = y + kDy[dir];
      if (!OnBoard(nx, ny)) continue;
      if (board_[nx + ny * d_].player == cur_) {
        alreadyjoined |= Join((int)a, nx + ny * d_);
        skip = true;          // the next neighbour touches this one: same group already, a sharp corner that cannot close a ring
      }
    }
    const Cell& g = board_[Leader((int)a)];
    if (Bits(g.edge) >= 3 || Bits(g.corner) >= 2 || (alreadyjoined && Ring(x, y, 0, 3))) outcome_ = cur_;
    else if (moves_ == valid_) outcome_ = kDraw;
    cur_ = 1 - cur_;
  }

 private:
  struct Cell { int player; bool mark; int parent, size, corner, edge; };
  static int Bits(int v) { int n = 0; for (; v; v &= v - 1) ++n; return n; }
  bool OnBoard(int x, int y) const { return x >= 0 && y >= 0 && x < d_ && y < d_ && y - x < size_ && x - y < size_; }   // havannah.h:58-66
  int Corner(int x, int y) const {                                                             // havannah.cc:128-142
    const int m = size_ - 1, e = 2 * m;
    if (x == 0 && y == 0) return 1;
    if (x == m && y == 0) return 2;
    if (x == e && y == m) return 4;
    if (x == e && y == e) return 8;
    if (x == m && y == e) return 16;
    if (x == 0 && y == m) return 32;
    return 0;
  }
  int Edge(int x, int y) const {                                                               // havannah.cc:144-158
    const int m = size_ - 1, e = 2 * m;
    if (y == 0 && x != 0 && x != m) return 1;
    if (x - y == m && x != m && x != e) return 2;
    if (x == e && y != m && y != e) return 4;
    if (y == e && x != e && x != m) return 8;
    if (y - x == m && x != m && x != 0) return 16;
    if (x == 0 && y != m && y != 0) return 32;
    return 0;
  }
  bool AllowSwap() const { return swap_ && moves_ == 1 && cur_ == kP2; }                      // havannah.cc:208-210
  int Leader(int c) { while (board_[c].parent != c) c = board_[c].parent; return c; }
  bool Join(int a, int b) {                                                                    // havannah.cc:375-392
    int la = Leader(a), lb = Leader(b);
    if (la == lb) return true;
    if (board_[la].size < board_[lb].size) std::swap(la, lb);
    board_[lb].parent = la;
    board_[la].size += board_[lb].size;
    board_[la].corner |= board_[lb].corner;
    board_[la].edge |= board_[lb].edge;
    return false;
  }
  bool Ring(int x, int y, int left, int right) {                                               // havannah.cc:394-409
    if (!OnBoard(x, y)) return false;
    Cell& c = board_[x + y * d_];
    if (c.player != cur_) return false;
    if (c.mark) return true;
    c.mark = true;
    bool success = false;
    for (int i = left; !success && i <= right; ++i) {
      const int dir = (i + 6) % 6;
      success = Ring(x + kDx[dir], y + kDy[dir], dir - 1, dir + 1);
    }
    c.mark = false;
    return success;
  }
  int size_, d_, valid_;
  bool swap_;
  std::vector<Cell> board_;
  int cur_ = kP1, outcome_ = kNone, moves_ = 0, last_ = -1;
};

class HavannahGame : public Game {
 public:
  explicit HavannahGame(const Params& p) {
    size_ = (int)p.get("board_size", 8);                 // havannah.h:38
    swap_ = p.get("swap", 0) != 0;
    const int d = 2 * size_ - 1;
    info.name = "havannah";
    info.num_distinct_actions = d * d;
    info.max_game_length = d * d - size_ * (size_ - 1) + (swap_ ? 1 : 0);
    info.observation_tensor_size = 3 * d * d;
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<HavannahState>(size_, swap_); }
 private:
  int size_;
  bool swap_;
};

}  // namespace
std::unique_ptr<Game> MakeHavannah(const Params& p) { return std::make_unique<HavannahGame>(p); }
}  // namespace oracle
