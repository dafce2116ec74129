#include "b200_games.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "open_spiel/games/breakthrough/breakthrough.h"
#include "open_spiel/games/connect_four/connect_four.h"
#include "open_spiel/games/go/go.h"
#include "open_spiel/games/hex/hex.h"
#include "open_spiel/games/kuhn_poker/kuhn_poker.h"
#include "open_spiel/games/leduc_poker/leduc_poker.h"
#include "open_spiel/games/tic_tac_toe/tic_tac_toe.h"

namespace open_spiel {
namespace b200 {
namespace {

void Check(int rc) {
  if (rc != 0) SpielFatalError(std::string("b2s: ") + b2s_last_error());   // the reference's own error path
}

std::string Num(double v) {        // absl::StrCat(double) / ostream << float print like "%g"
  char b[64];
  snprintf(b, sizeof b, "%g", v);
  return b;
}

constexpr int kLeducInvalidCard = -10000;   // leduc_poker.h:60
constexpr int kLeducStartingMoney = 100;    // leduc_poker.h:68

// pot_ / money_ of LeducState (leduc_poker.cc:628-678, 702-706) from the antes: the pot is paid out at the terminal state
void LeducMoney(const b2s_host::Decoded& d, bool terminal, const float* returns, int* pot, double money[2]) {
  *pot = d.ante[0] + d.ante[1];
  for (int p = 0; p < 2; ++p) money[p] = kLeducStartingMoney - d.ante[p];
  if (terminal) {
    for (int p = 0; p < 2; ++p) money[p] = kLeducStartingMoney + (double)returns[p];   // Returns = money - starting money
    *pot = 0;
  }
}

}  // namespace

// ---- game -------------------------------------------------------------------------------------------------------

std::shared_ptr<const Game> B200Game::Create(const GameType& type, const GameParameters& params) {
  std::shared_ptr<B200Game> g(new B200Game(type, params));
  const std::string& name = type.short_name;
  g->gid_ = b2s_game_id(name.c_str());
  if (g->gid_ < 0) return nullptr;
  b2s_params& p = g->cparams_;
  b2s_params_default(&p);
  // ParameterValue<> records the defaults used, so GetParameters() / ToString() print what the stock game prints
  // (the calls mirror the stock constructors: connect_four.cc:333-340, breakthrough.cc:383-386, hex.cc:404-414,
  // go.cc:303-309, kuhn_poker.cc:375-376, leduc_poker.cc:782-788).
  if (name == "connect_four") {
    p.egocentric_obs_tensor = g->ParameterValue<bool>("egocentric_obs_tensor") ? 1 : 0;
    p.rows = g->ParameterValue<int>("rows");
    p.columns = g->ParameterValue<int>("columns");
    p.x_in_row = g->ParameterValue<int>("x_in_row");
  } else if (name == "breakthrough") {
    p.rows = g->ParameterValue<int>("rows");
    p.columns = g->ParameterValue<int>("columns");
  } else if (name == "hex") {
    p.columns = g->ParameterValue<int>("num_cols", g->ParameterValue<int>("board_size"));
    p.rows = g->ParameterValue<int>("num_rows", g->ParameterValue<int>("board_size"));
    const std::string rep = g->ParameterValue<std::string>("string_rep", "standard");
    if (rep != "standard" && rep != "explicit") SpielFatalError("Invalid string_rep " + rep);   // hex.cc:66-73
    g->hex_explicit_ = rep == "explicit";
    p.swap = g->ParameterValue<bool>("swap") ? 1 : 0;
    p.plain_obs_tensor = g->ParameterValue<bool>("plain_obs_tensor") ? 1 : 0;
  } else if (name == "go") {
    p.komi = g->ParameterValue<double>("komi");
    p.board_size = g->ParameterValue<int>("board_size");
    p.handicap = g->ParameterValue<int>("handicap");
    if (p.board_size < 1 || p.board_size > 19) return nullptr;
    p.max_game_length = g->ParameterValue<int>("max_game_length", p.board_size * p.board_size * 2);   // go.h:68-70
    g->komi_ = (float)p.komi;
  } else if (name == "kuhn_poker") {
    p.players = g->ParameterValue<int>("players");
  } else if (name == "leduc_poker") {
    p.players = g->ParameterValue<int>("players");
    const bool action_mapping = g->ParameterValue<bool>("action_mapping");
    const bool suit_isomorphism = g->ParameterValue<bool>("suit_isomorphism");
    p.starting_player = g->ParameterValue<int>("starting_player");
    if (action_mapping || suit_isomorphism) return nullptr;      // not representable in the packed layout
  }
  std::string err;
  g->rules_ = b2s_host::Rules::Create(g->gid_, p, &err);
  if (!g->rules_) return nullptr;
  return g;
}

std::unique_ptr<State> B200Game::NewInitialState() const {
  return std::unique_ptr<State>(new B200State(shared_from_this()));
}

std::vector<int> B200Game::ObservationTensorShape() const {
  std::vector<int> s;
  for (int d : info().obs_shape) if (d > 0) s.push_back(d);
  return s;
}

std::vector<int> B200Game::InformationStateTensorShape() const {
  if (info().information_state_tensor_size <= 0) return Game::InformationStateTensorShape();
  return {info().information_state_tensor_size};
}

void* B200Game::NewBatch(int64_t n, int device) const {
  void* b = nullptr;
  Check(b2s_batch_create(gid_, &cparams_, n, device, &b));
  return b;
}

// Action strings that do not depend on the state (tic_tac_toe.cc:266-270, leduc_poker.cc:864-870 and the
// State::ActionToString of the other games, which only read game parameters).
std::string B200Game::ActionToString(Player player, Action a) const {
  const b2s_game_info& gi = info();
  switch (gid_) {
    case B2S_TIC_TAC_TOE:
      return std::string(player == 0 ? "x" : "o") + "(" + std::to_string(a / 3) + "," + std::to_string(a % 3) + ")";
    case B2S_CONNECT_FOUR:
      return std::string(player == 0 ? "x" : "o") + std::to_string(a);
    case B2S_BREAKTHROUGH: {      // breakthrough.cc:196-217: from-cell, to-cell, '*' for captures
      const int rows = gi.obs_shape[1], cols = gi.obs_shape[2];
      const int cap = (int)(a & 1), dir = (int)((a >> 1) % 6), cell = (int)(a / 12);
      const int r1 = cell / cols, c1 = cell % cols;
      const int r2 = dir < 3 ? r1 + 1 : r1 - 1, c2 = c1 + dir % 3 - 1;
      std::string s;
      s += (char)('a' + c1); s += (char)('1' + (rows - 1 - r1));
      s += (char)('a' + c2); s += (char)('1' + (rows - 1 - r2));
      if (cap) s += "*";
      return s;
    }
    case B2S_HEX: {               // hex.cc:295-314 (standard representation)
      const int cols = gi.obs_shape[1];
      if (cparams_.swap > 0 && a == gi.num_distinct_actions - 1) return "swap";
      const int row = (int)(a % cols), col = (int)(a / cols);
      std::string s(1, (char)('a' + row));
      return s + std::to_string(col + 1);
    }
    case B2S_GO: {                // go.cc:172-176, go_board.cc:229-242
      const int n = gi.obs_shape[1];
      std::string s = player == 0 ? "B " : "W ";
      if (a == (Action)n * n) return s + "PASS";
      char col = (char)('a' + a % n);
      if (col >= 'i') ++col;      // Go / SGF labelling skips 'i'
      return s + std::string(1, col) + std::to_string(a / n + 1);
    }
    case B2S_KUHN_POKER:          // kuhn_poker.cc:244-251
      if (player == kChancePlayerId) return "Deal:" + std::to_string(a);
      return a == 0 ? "Pass" : "Bet";
    case B2S_LEDUC_POKER:         // leduc_poker.cc:864-870, 67-78
      if (player == kChancePlayerId) return "Chance outcome:" + std::to_string(a);
      if (a == 0) return "Fold";
      if (a == 1) return "Call";
      if (a == 2) return "Raise";
      SpielFatalError("Unknown action: " + std::to_string(a));
  }
  return std::to_string(a);
}

// ---- state ------------------------------------------------------------------------------------------------------

B200State::B200State(std::shared_ptr<const Game> game) : State(game) {
  blob_.resize((rules().blob_bytes() + sizeof(Word16) - 1) / sizeof(Word16));
  rules().Init(blob_.data());
}

Player B200State::CurrentPlayer() const { return rules().CurrentPlayer(blob_.data()); }
bool B200State::IsTerminal() const { return rules().CurrentPlayer(blob_.data()) == kTerminalPlayerId; }

std::vector<double> B200State::Returns() const {
  float r[8];
  rules().Returns(blob_.data(), r);
  return std::vector<double>(r, r + num_players_);        // float -> double is exact for every value the games produce
}

std::vector<Action> B200State::LegalActions() const {
  uint32_t m[32];
  const int words = rules().info().mask_words;
  rules().LegalMask(blob_.data(), m);
  std::vector<Action> out;                                  // ascending ids, empty at terminal states (spiel.h:374-388)
  for (int w = 0; w < words; ++w)
    for (uint32_t bits = m[w]; bits; bits &= bits - 1) out.push_back(w * 32 + __builtin_ctz(bits));
  return out;
}

std::vector<std::pair<Action, double>> B200State::ChanceOutcomes() const {
  SPIEL_CHECK_TRUE(IsChanceNode());
  // kuhn_poker.cc:329-337, leduc_poker.cc:546-571: uniform over the cards still in the deck
  std::vector<Action> cards = LegalActions();
  std::vector<std::pair<Action, double>> out;
  const double p = 1.0 / (double)cards.size();
  for (Action c : cards) out.push_back({c, p});
  return out;
}

void B200State::DoApplyAction(Action action_id) {
  const size_t sw = (rules().state_bytes() + sizeof(Word16) - 1) / sizeof(Word16);
  undo_.insert(undo_.end(), blob_.begin(), blob_.begin() + sw);
  if (!rules().Apply(blob_.data(), (int)action_id)) {
    undo_.resize(undo_.size() - sw);
    // the stock games SPIEL_CHECK inside DoApplyAction (e.g. connect_four.cc:131-133)
    SpielFatalError("b200: illegal action " + std::to_string(action_id) + " in state\n" + ToString());
  }
}

void B200State::UndoAction(Player player, Action action) {
  const size_t sw = (rules().state_bytes() + sizeof(Word16) - 1) / sizeof(Word16);
  SPIEL_CHECK_GE(undo_.size(), sw);
  SPIEL_CHECK_FALSE(history_.empty());
  SPIEL_CHECK_EQ(history_.back().action, action);
  std::copy(undo_.end() - sw, undo_.end(), blob_.begin());
  undo_.resize(undo_.size() - sw);
  history_.pop_back();
  --move_number_;
}

void B200State::ObservationTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  SPIEL_CHECK_EQ((int)values.size(), rules().info().observation_tensor_size);
  SPIEL_CHECK_TRUE(rules().Tensor(blob_.data(), player, 0, values.data()));
}

void B200State::InformationStateTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  SPIEL_CHECK_EQ((int)values.size(), rules().info().information_state_tensor_size);
  if (!rules().Tensor(blob_.data(), player, 1, values.data())) SpielFatalError("InformationStateTensor unimplemented!");
}

std::unique_ptr<State> B200State::Clone() const { return std::unique_ptr<State>(new B200State(*this)); }

void B200State::ToBatchLane(void* batch, int64_t lane) const { Check(b2s_state_set(batch, lane, blob_.data(), rules().blob_bytes())); }
void B200State::FromBatchLane(void* batch, int64_t lane) { Check(b2s_state_get(batch, lane, blob_.data(), rules().blob_bytes())); }

std::string B200State::ActionToString(Player player, Action action_id) const {
  // hex.cc:301 tests `StringRep() == StringRep::kStandard` on a value-initialised enum, not on the state's string_rep_,
  // so the stock game prints the standard form whatever string_rep says; all seven games depend on parameters only.
  return bgame().ActionToString(player, action_id);
}

// Reference formats: tic_tac_toe.cc:165-176, connect_four.cc:212-222, breakthrough.cc:264-284, hex.cc:343-359,
// go.cc:178-184 + go_board.cc:566-583, kuhn_poker.cc:253-268, leduc_poker.cc:463-496.
std::string B200State::ToString() const {
  const b2s_game_info& gi = rules().info();
  b2s_host::Decoded d;
  rules().Decode(blob_.data(), &d);
  std::string s;
  switch (bgame().gid()) {
    case B2S_TIC_TAC_TOE:
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) s += ".xo"[d.cells[r * 3 + c]];
        if (r < 2) s += "\n";
      }
      return s;
    case B2S_CONNECT_FOUR: {
      const int rows = gi.obs_shape[1], cols = gi.obs_shape[2];
      for (int r = rows - 1; r >= 0; --r) {
        for (int c = 0; c < cols; ++c) s += ".xo"[d.cells[r * cols + c]];
        s += "\n";
      }
      return s;
    }
    case B2S_BREAKTHROUGH: {
      const int rows = gi.obs_shape[1], cols = gi.obs_shape[2];
      for (int r = 0; r < rows; ++r) {
        s += (char)('1' + (rows - 1 - r));
        for (int c = 0; c < cols; ++c) s += ".bw"[d.cells[r * cols + c]];
        s += "\n";
      }
      s += " ";
      for (int c = 0; c < cols; ++c) s += (char)('a' + c);
      s += "\n";
      return s;
    }
    case B2S_HEX: {
      const int cols = gi.obs_shape[1];
      const char* chars = bgame().hex_explicit() ? ".xyzXopqO" : ".xxxxoooo";
      int line = 0;
      for (int cell = 0; cell < (int)d.cells.size(); ++cell) {
        if (cell && cell % cols == 0) { s += "\n"; ++line; s += std::string(line, ' '); }
        s += chars[d.cells[cell]];
        s += " ";
      }
      return s;
    }
    case B2S_GO: {
      const int n = gi.obs_shape[1];
      s = "GoState(komi=" + Num(bgame().komi()) + ", to_play=" + (d.to_play == 0 ? "B" : "W") +
          ", history.size()=" + std::to_string(history_.size()) + ")\n\n";
      for (int row = n - 1; row >= 0; --row) {
        char b[8];
        snprintf(b, sizeof b, "%2d ", row + 1);
        s += b;
        for (int col = 0; col < n; ++col) s += "+XO"[d.cells[row * n + col]];
        s += "\n";
      }
      s += "   " + std::string("ABCDEFGHJKLMNOPQRST").substr(0, n) + "\n";
      return s;
    }
    case B2S_KUHN_POKER: {
      for (int i = 0; i < (int)history_.size() && i < num_players_; ++i) {
        if (!s.empty()) s += ' ';
        s += std::to_string(history_[i].action);
      }
      if ((int)history_.size() > num_players_) s += ' ';
      for (int i = num_players_; i < (int)history_.size(); ++i) s += history_[i].action ? 'b' : 'p';
      return s;
    }
    case B2S_LEDUC_POKER: {
      static const char* kNames[3] = {"Fold", "Call", "Raise"};
      float ret[2];
      rules().Returns(blob_.data(), ret);
      int pot;
      double money[2];
      LeducMoney(d, IsTerminal(), ret, &pot, money);
      s = "Round: " + std::to_string(d.round) + "\nPlayer: " + std::to_string(d.cur_player) + "\nPot: " + std::to_string(pot) +
          "\nMoney (player_0 player_1):";
      for (int p = 0; p < 2; ++p) s += " " + Num(money[p]);
      s += "\nCards (public player_0 player_1): " + std::to_string(d.public_card < 0 ? kLeducInvalidCard : d.public_card) + " ";
      for (int p = 0; p < 2; ++p) s += std::to_string(d.private_card[p] < 0 ? kLeducInvalidCard : d.private_card[p]) + " ";
      s += "\nRound 1 sequence: ";
      for (size_t i = 0; i < d.round1.size(); ++i) { if (i) s += ", "; s += kNames[d.round1[i]]; }
      s += "\nRound 2 sequence: ";
      for (size_t i = 0; i < d.round2.size(); ++i) { if (i) s += ", "; s += kNames[d.round2[i]]; }
      s += "\n";
      return s;
    }
  }
  return HistoryString();
}

// kuhn_poker.cc:109-166 (KuhnObserver::StringFrom), leduc_poker.cc:198-239 (LeducObserver::StringFrom); the board games
// return HistoryString() / ToString() (e.g. connect_four.cc:287-297).
std::string B200State::InformationStateString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  const int gid = bgame().gid();
  if (gid == B2S_KUHN_POKER) {
    std::string s;
    if ((int)history_.size() > player) s += std::to_string(history_[player].action);
    for (int i = num_players_; i < (int)history_.size(); ++i) s += history_[i].action ? 'b' : 'p';
    return s;
  }
  if (gid == B2S_LEDUC_POKER) {
    b2s_host::Decoded d;
    rules().Decode(blob_.data(), &d);
    float ret[2];
    rules().Returns(blob_.data(), ret);
    int pot;
    double money[2];
    LeducMoney(d, IsTerminal(), ret, &pot, money);
    auto join = [](const std::vector<int>& v) { std::string t; for (size_t i = 0; i < v.size(); ++i) { if (i) t += " "; t += std::to_string(v[i]); } return t; };
    std::string s = "[Observer: " + std::to_string(player) + "][Private: " +
                    std::to_string(d.private_card[player] < 0 ? kLeducInvalidCard : d.private_card[player]) + "]";
    s += "[Round " + std::to_string(d.round) + "][Player: " + std::to_string(d.cur_player) + "][Pot: " + std::to_string(pot) +
         "][Money: " + Num(money[0]) + " " + Num(money[1]) + "]";
    if (d.public_card >= 0) s += "[Public: " + std::to_string(d.public_card) + "]";
    s += "[Round1: " + join(d.round1) + "][Round2: " + join(d.round2) + "]";
    return s;
  }
  return HistoryString();
}

std::string B200State::ObservationString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  const int gid = bgame().gid();
  if (gid == B2S_KUHN_POKER) {
    std::string s;
    if ((int)history_.size() > player) {
      s += std::to_string(history_[player].action);
      float obs[16];
      rules().Tensor(blob_.data(), player, 0, obs);            // pot contributions are the last num_players_ entries
      const int off = rules().info().observation_tensor_size - num_players_;
      for (int p = 0; p < num_players_; ++p) s += std::to_string((int)obs[off + p]);
    }
    return s;
  }
  if (gid == B2S_LEDUC_POKER) {
    b2s_host::Decoded d;
    rules().Decode(blob_.data(), &d);
    float ret[2];
    rules().Returns(blob_.data(), ret);
    int pot;
    double money[2];
    LeducMoney(d, IsTerminal(), ret, &pot, money);
    std::string s = "[Observer: " + std::to_string(player) + "][Private: " +
                    std::to_string(d.private_card[player] < 0 ? kLeducInvalidCard : d.private_card[player]) + "]";
    s += "[Round " + std::to_string(d.round) + "][Player: " + std::to_string(d.cur_player) + "][Pot: " + std::to_string(pot) +
         "][Money: " + Num(money[0]) + " " + Num(money[1]) + "]";
    if (d.public_card >= 0) s += "[Public: " + std::to_string(d.public_card) + "]";
    s += "[Ante: " + std::to_string(d.ante[0]) + " " + std::to_string(d.ante[1]) + "]";
    return s;
  }
  return ToString();
}

// ---- registration -------------------------------------------------------------------------------------------------

namespace {
// The stock game for parameter sets the packed layouts cannot hold (the factory of the stock registration is private
// to GameRegisterer, so the stock Game classes are constructed directly; their constructors are public).
std::shared_ptr<const Game> StockGame(const std::string& name, const GameParameters& params) {
  if (name == "tic_tac_toe") return std::shared_ptr<const Game>(new tic_tac_toe::TicTacToeGame(params));
  if (name == "connect_four") return std::shared_ptr<const Game>(new connect_four::ConnectFourGame(params));
  if (name == "breakthrough") return std::shared_ptr<const Game>(new breakthrough::BreakthroughGame(params));
  if (name == "hex") return std::shared_ptr<const Game>(new hex::HexGame(params));
  if (name == "go") return std::shared_ptr<const Game>(new go::GoGame(params));
  if (name == "kuhn_poker") return std::shared_ptr<const Game>(new kuhn_poker::KuhnGame(params));
  if (name == "leduc_poker") return std::shared_ptr<const Game>(new leduc_poker::LeducGame(params));
  SpielFatalError("b200: no stock game " + name);
}
}  // namespace

void RegisterB200Games() {
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"tic_tac_toe", "connect_four", "breakthrough", "hex", "go", "kuhn_poker", "leduc_poker"}) {
      GameType type;
      for (const GameType& t : GameRegisterer::RegisteredGames())
        if (t.short_name == name) type = t;                    // the stock registration's GameType, unchanged
      SPIEL_CHECK_EQ(type.short_name, std::string(name));
      GameRegisterer::RegisterGame(type, [type](const GameParameters& params) {
        std::shared_ptr<const Game> g = B200Game::Create(type, params);
        if (g) return g;
        return StockGame(type.short_name, params);
      });
    }
  });
}

}  // namespace b200
}  // namespace open_spiel
