// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// CPU restatement of reference open_spiel/games/go/{go,go_board}.{h,cc}.
// Keeps the reference's representation: 21x21 guarded "virtual" board, circular linked-list chains
// with pseudo-liberty sums (sum / sum-of-squares atari test), Zobrist positional superko set,
// recursive Tromp-Taylor flood.  The CUDA kernels use 81-bit bitboards + flood fill instead.
#include <array>
#include <cstdio>
#include <functional>
#include <random>
#include <unordered_set>

#include "../oracle.h"

namespace oracle {
namespace {

constexpr int kVB = 21;                 // go_board.h:49-52
constexpr int kVPoints = kVB * kVB;
constexpr int kVPass = kVPoints + 1;    // go_board.h:57
enum Color : uint8_t { kBlack = 0, kWhite = 1, kEmpty = 2, kGuard = 3 };   // go_board.h:30

// chess_common.h:129-170 — ZobristTable<uint64_t, 441, 2>(seed): an outer mt19937_64 draws one seed
// per point; each inner mt19937_64 draws the two colour values.  The reference draws through
// absl::uniform_int_distribution<uint64_t>() whose full-range case returns the engine word unchanged
// (abseil-cpp 20250814.1, absl/random/uniform_int_distribution.h: range == max() branch).
const std::array<std::array<uint64_t, 2>, kVPoints>& Zobrist() {
  static const auto table = [] {
    std::array<std::array<uint64_t, 2>, kVPoints> t;
    std::mt19937_64 outer(2765481);     // go_board.cc:356-358
    for (int i = 0; i < kVPoints; ++i) {
      std::mt19937_64 inner(outer());
      t[i][0] = inner();
      t[i][1] = inner();
    }
    return t;
  }();
  return table;
}

Color Opp(Color c) { return c == kBlack ? kWhite : c == kWhite ? kBlack : c; }

struct Chain {                          // go_board.h:233-257
  uint32_t sum_sq;
  uint16_t sum, stones, libs;
  void reset() { sum_sq = 0; sum = 0; stones = 0; libs = 0; }
  void reset_border() { stones = 0; libs = 4; sum = 32768; sum_sq = 2147483648u; }
  void merge(const Chain& o) { stones += o.stones; libs += o.libs; sum += o.sum; sum_sq += o.sum_sq; }
  bool in_atari() const { return (uint32_t)libs * sum_sq == (uint32_t)sum * (uint32_t)sum; }
  void add(int p) { libs += 1; sum += p; sum_sq += (uint32_t)p * (uint32_t)p; }
  void remove(int p) { libs -= 1; sum -= p; sum_sq -= (uint32_t)p * (uint32_t)p; }
};

class Board {
 public:
  explicit Board(int n) : n_(n) { Clear(); }
  int size() const { return n_; }
  static int VPoint(int row, int col) { return (row + 1) * kVB + col + 1; }   // go_board.cc:134-137
  int ToVirtual(int64_t a) const { return a == n_ * n_ ? kVPass : VPoint((int)(a / n_), (int)(a % n_)); }
  Color color(int p) const { return col_[p]; }
  uint64_t hash() const { return hash_; }
  int ko() const { return ko_; }

  void Clear() {                        // go_board.cc:269-297
    hash_ = 0;
    for (int i = 0; i < kVPoints; ++i) { col_[i] = kGuard; head_[i] = next_[i] = i; ch_[i].reset_border(); }
    ForBoard([&](int p) { col_[p] = kEmpty; ch_[p].reset(); });
    ForBoard([&](int p) { Nb(p, [&](int q) { if (col_[q] == kEmpty) chain(p).add(q); }); });
    ko_ = 0;
  }
  bool InArea(int p) const {
    int r = p / kVB - 1, c = p % kVB - 1;
    return p != 0 && p != kVPass && r >= 0 && r < n_ && c >= 0 && c < n_;
  }
  // go_board.cc:481-506
  bool IsLegal(int p, Color c) const {
    if (p == kVPass) return true;
    if (!InArea(p)) return false;
    if (col_[p] != kEmpty || p == ko_) return false;
    if (chain(p).libs > 0) return true;
    bool ok = false;
    Nb(p, [&](int q) { ok |= (col_[q] == c && !chain(q).in_atari()); });
    if (ok) return true;
    Nb(p, [&](int q) { ok |= (col_[q] == Opp(c) && chain(q).in_atari()); });
    return ok;
  }
  // go_board.cc:299-336
  bool Play(int p, Color c) {
    if (p == kVPass) { ko_ = 0; return true; }
    if (col_[p] != kEmpty) return false;
    bool in_enemy_eye = true;
    Nb(p, [&](int q) { if (col_[q] == c || col_[q] == kEmpty) in_enemy_eye = false; });
    Join(p, c);
    SetStone(p, c);
    Nb(p, [&](int q) { chain(q).remove(p); });
    int first_capture = 0, captured = 0;
    bool have_first = false;
    Nb(p, [&](int q) {                  // CaptureDeadChains, go_board.cc:423-439
      if (col_[q] == Opp(c) && chain(q).libs == 0) {
        if (!have_first) { first_capture = head_[q]; have_first = true; }
        captured += chain(q).stones;
        RemoveChain(q);
      }
    });
    ko_ = (in_enemy_eye && captured == 1) ? first_capture : 0;
    return true;
  }
  template <typename F> void ForBoard(F f) const {
    for (int r = 0; r < n_; ++r) for (int c = 0; c < n_; ++c) f(VPoint(r, c));
  }
  template <typename F> static void Nb(int p, F f) { f(p + kVB); f(p + 1); f(p - 1); f(p - kVB); }  // go_board.cc:59-66

 private:
  Chain& chain(int p) { return ch_[head_[p]]; }
  const Chain& chain(int p) const { return ch_[head_[p]]; }
  void SetStone(int p, Color c) {       // go_board.cc:355-364
    hash_ ^= Zobrist()[p][c == kEmpty ? col_[p] : c];
    col_[p] = c;
  }
  void InitChain(int p) {               // go_board.cc:460-472
    head_[p] = next_[p] = p;
    Chain& c = ch_[p];
    c.reset();
    c.stones += 1;
    Nb(p, [&](int q) { if (col_[q] == kEmpty) c.add(q); });
  }
  void Join(int p, Color c) {           // go_board.cc:368-417
    int big = 0, big_size = 0;
    Nb(p, [&](int q) {
      if (col_[q] == c && chain(q).stones > big_size) { big_size = chain(q).stones; big = head_[q]; }
    });
    if (big_size == 0) { InitChain(p); return; }
    Nb(p, [&](int q) {
      if (col_[q] == c && head_[q] != big) {
        ch_[big].merge(chain(q));
        int cur = q;
        do { head_[cur] = big; cur = next_[cur]; } while (cur != q);
        std::swap(next_[big], next_[q]);
      }
    });
    next_[p] = next_[big];
    next_[big] = p;
    head_[p] = big;
    ch_[big].stones += 1;
    Nb(p, [&](int q) { if (col_[q] == kEmpty) ch_[big].add(q); });
  }
  void RemoveChain(int p) {             // go_board.cc:441-458
    int this_head = head_[p], cur = p;
    do {
      int nxt = next_[cur];
      SetStone(cur, kEmpty);
      InitChain(cur);
      Nb(cur, [&](int q) { if (head_[q] != this_head || col_[q] == kEmpty) chain(q).add(cur); });
      cur = nxt;
    } while (cur != p);
  }
  int n_;
  Color col_[kVPoints];
  uint16_t head_[kVPoints], next_[kVPoints];
  Chain ch_[kVPoints];
  uint64_t hash_ = 0;
  int ko_ = 0;
};

// go_board.cc:612-683
int Surrounded(const Board& b, int p, std::array<bool, kVPoints>& marked, bool& rb, bool& rw) {
  if (marked[p]) return 0;
  marked[p] = true;
  int n = 1;
  Board::Nb(p, [&](int q) {
    switch (b.color(q)) {
      case kBlack: rb = true; break;
      case kWhite: rw = true; break;
      case kEmpty: n += Surrounded(b, q, marked, rb, rw); break;
      default: break;
    }
  });
  return n;
}
float TrompTaylor(const Board& b, float komi, int handicap) {
  int delta = 0;
  std::array<bool, kVPoints> marked;
  marked.fill(false);
  b.ForBoard([&](int p) {
    if (b.color(p) == kBlack) ++delta;
    else if (b.color(p) == kWhite) --delta;
    else if (!marked[p]) {
      bool rb = false, rw = false;
      int n = Surrounded(b, p, marked, rb, rw);
      if (rb && !rw) delta += n;
      else if (!rb && rw) delta -= n;
    }
  });
  float score = delta - komi;
  if (handicap >= 2) score -= handicap;
  return score;
}

struct GoCfg { int n; float komi; int handicap; int max_len; };

class GoState : public State {
 public:
  explicit GoState(const GoCfg& g) : g_(g), board_(g.n) { Reset(); }
  int CurrentPlayer() const override { return IsTerminal() ? kTerminalPlayerId : (int)to_play_; }   // go.h:90-92

  // go.cc:160-170
  std::vector<int64_t> LegalActions() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int r = 0; r < g_.n; ++r)
      for (int c = 0; c < g_.n; ++c)
        if (board_.IsLegal(Board::VPoint(r, c), to_play_)) v.push_back(r * g_.n + c);
    v.push_back(g_.n * g_.n);
    return v;
  }
  std::vector<int64_t> RolloutCandidates() const override {
    std::vector<int64_t> v;
    if (IsTerminal()) return v;
    for (int r = 0; r < g_.n; ++r)
      for (int c = 0; c < g_.n; ++c) {
        int p = Board::VPoint(r, c);
        if (board_.color(p) == kEmpty && p != board_.ko()) v.push_back(r * g_.n + c);
      }
    v.push_back(g_.n * g_.n);
    return v;
  }
  // go.cc:225-230
  bool IsTerminal() const override {
    size_t h = history_.size();
    if (h < 2) return false;
    int pass = g_.n * g_.n;
    return (int)h >= g_.max_len || superko_ || (history_[h - 1].second == pass && history_[h - 2].second == pass);
  }
  // go.cc:232-258
  std::vector<double> Returns() const override {
    if (!IsTerminal()) return {0.0, 0.0};
    if (superko_) return {0.0, 0.0};
    float s = TrompTaylor(board_, g_.komi, g_.handicap);
    if (s > 0) return {1.0, -1.0};
    if (s < 0) return {-1.0, 1.0};
    return {0.0, 0.0};
  }
  // go.cc:178-184, go_board.cc:566-596
  std::string ToString() const override {
    // (snprintf instead of iostreams: ostream << float crashes when torch's libstdc++ is co-loaded)
    char b[96];
    snprintf(b, sizeof b, "GoState(komi=%g, to_play=%s, history.size()=%zu)\n\n", (double)g_.komi,
             to_play_ == kBlack ? "B" : "W", history_.size());
    std::string ss = b;
    for (int row = g_.n - 1; row >= 0; --row) {
      snprintf(b, sizeof b, "%2d ", row + 1);
      ss += b;
      for (int col = 0; col < g_.n; ++col) ss += "XO+#"[board_.color(Board::VPoint(row, col))];
      ss += "\n";
    }
    ss += "   " + std::string("ABCDEFGHJKLMNOPQRST").substr(0, g_.n) + "\n";
    return ss;
  }
  // go.cc:138-158
  void ObservationTensor(int, float* out) const override {
    int n = g_.n * g_.n;
    for (int i = 0; i < 4 * n; ++i) out[i] = 0.f;
    int cell = 0;
    board_.ForBoard([&](int p) { out[n * (int)board_.color(p) + cell] = 1.f; ++cell; });
    for (int i = 0; i < n; ++i) out[3 * n + i] = to_play_ == kWhite ? 1.f : 0.f;
  }
  std::string InformationStateString(int) const override {
    std::string s;
    for (size_t i = 0; i < history_.size(); ++i) { if (i) s += ", "; s += std::to_string(history_[i].second); }
    return s;
  }
  std::unique_ptr<State> Clone() const override { return std::make_unique<GoState>(*this); }
  uint64_t Hash() const { return board_.hash(); }

 protected:
  // go.cc:275-285
  void DoApplyAction(int64_t a) override {
    if (a < 0 || a > g_.n * g_.n || !board_.Play(board_.ToVirtual(a), to_play_)) { Fail("go: point occupied"); return; }
    to_play_ = Opp(to_play_);
    bool inserted = seen_.insert(board_.hash()).second;
    if (!inserted && a != g_.n * g_.n) superko_ = true;
  }

 private:
  // go.cc:287-301 (handicap stones, go.cc:72-93, use 19x19 coordinates d4,q16,... regardless of size)
  void Reset() {
    board_.Clear();
    if (g_.handicap < 2) {
      to_play_ = kBlack;
    } else {
      static const int pts[9][2] = {{3, 3}, {15, 15}, {15, 3}, {3, 15}, {9, 3}, {9, 15}, {3, 9}, {15, 9}, {9, 9}};
      int h = g_.handicap > 9 ? 0 : g_.handicap;
      for (int i = 0; i < h; ++i) {
        int r = pts[i][0], c = pts[i][1];
        if (h >= 5 && h % 2 == 1 && i == h - 1) { r = 9; c = 9; }
        board_.Play(Board::VPoint(r, c), kBlack);
      }
      to_play_ = kWhite;
    }
    seen_.clear();
    seen_.insert(board_.hash());
    superko_ = false;
  }
  GoCfg g_;
  Board board_;
  Color to_play_ = kBlack;
  bool superko_ = false;
  std::unordered_set<uint64_t> seen_;
};

class GoGame : public Game {
 public:
  explicit GoGame(const Params& p) {
    cfg_.n = (int)p.get("board_size", 19);                 // go.h:47-49
    cfg_.komi = (float)p.get("komi", 7.5);
    cfg_.handicap = (int)p.get("handicap", 0);
    cfg_.max_len = (int)p.get("max_game_length", cfg_.n * cfg_.n * 2);   // go.h:68-70
    info.name = "go";
    info.num_distinct_actions = cfg_.n * cfg_.n + 1;       // go.h:61-63
    info.max_game_length = cfg_.max_len;
    info.observation_tensor_size = 4 * cfg_.n * cfg_.n;    // go.h:175-179
  }
  std::unique_ptr<State> NewInitialState() const override { return std::make_unique<GoState>(cfg_); }
 private:
  GoCfg cfg_;
};

}  // namespace
std::unique_ptr<Game> MakeGo(const Params& p) { return std::make_unique<GoGame>(p); }
}  // namespace oracle
