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
#include "open_spiel/games/mnk/mnk.h"
#include "open_spiel/games/othello/othello.h"
// y.h defines the free function CalcXY without `inline` (y.h:57-64), so a second translation unit that includes it collides
// with y.o at link time; the name is redirected for this include (the function body is the header's own).
#define CalcXY CalcXY_b200_adapter
#include "open_spiel/games/y/y.h"
#undef CalcXY
#include "open_spiel/games/havannah/havannah.h"
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
  } else if (name == "mnk") {
    // read without recording: see TouchBoardParams() in the header
    auto raw = [&](const char* key, int def) {
      auto it = params.find(key);
      return it == params.end() ? def : it->second.int_value();
    };
    p.columns = raw("m", 15);      // mnk.h:34-36
    p.rows = raw("n", 15);
    p.x_in_row = raw("k", 5);
  } else if (name == "havannah") {
    p.board_size = g->ParameterValue<int>("board_size");              // havannah.cc:425-429
    if (g->ParameterValue<bool>("ansi_color_output")) return nullptr;  // escape sequences in ToString: the stock class prints them
    p.swap = g->ParameterValue<bool>("swap") ? 1 : 0;
  } else if (name == "y") {
    p.board_size = g->ParameterValue<int>("board_size");              // y.cc:331-334
    if (g->ParameterValue<bool>("ansi_color_output")) return nullptr;  // escape sequences in ToString: the stock class prints them
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
  TouchBoardParams();            // MNKState's constructor sizes the board (mnk.cc:172-178)
  return std::unique_ptr<State>(new B200State(shared_from_this()));
}

std::vector<int> B200Game::ObservationTensorShape() const {
  TouchBoardParams();
  std::vector<int> s;
  for (int d : info().obs_shape) if (d > 0) s.push_back(d);
  return s;
}

std::vector<int> B200Game::InformationStateTensorShape() const {
  if (info().information_state_tensor_size <= 0) return Game::InformationStateTensorShape();
  return {info().information_state_tensor_size};
}

// Structured observers of the two poker games, written against B200State::Poker().  Field names, shapes, order and the
// string formats are the stock observers' (kuhn_poker.cc:64-170, leduc_poker.cc:92-239), so the named tensors and the
// public / private observation strings of a playthrough come out identical.
namespace {
class B200PokerObserver : public Observer {
 public:
  B200PokerObserver(int gid, IIGObservationType t) : Observer(/*has_string=*/true, /*has_tensor=*/true), gid_(gid), t_(t) {}

  void WriteTensor(const State& observed_state, int player, Allocator* allocator) const override {
    const B200State& state = open_spiel::down_cast<const B200State&>(observed_state);
    const B200State::PokerView v = state.Poker();
    SPIEL_CHECK_GE(player, 0);
    SPIEL_CHECK_LT(player, v.num_players);
    const int n = v.num_players;
    const bool single = t_.private_info == PrivateInfoType::kSinglePlayer;
    if (gid_ == B2S_KUHN_POKER) {
      if (single) {
        allocator->Get("player", {n}).at(player) = 1;
        auto card = allocator->Get("private_card", {n + 1});
        if (v.private_cards[player] >= 0) card.at(v.private_cards[player]) = 1;
      }
      if (t_.public_info) {
        if (t_.perfect_recall) {
          auto out = allocator->Get("betting", {2 * n - 1, 2});
          for (size_t i = 0; i < v.round1.size(); ++i) out.at((int)i, v.round1[i]) = 1;
        } else {
          auto out = allocator->Get("pot_contribution", {n});
          for (int p = 0; p < n; ++p) out.at(p) = v.ante[p];
        }
      }
      return;
    }
    const int cards = 2 * (n + 1);                       // leduc: NumObservableCards without suit isomorphism
    allocator->Get("player", {n}).at(player) = 1;
    if (single) {
      auto out = allocator->Get("private_card", {cards});
      if (v.private_cards[player] >= 0) out.at(v.private_cards[player]) = 1;
    } else if (t_.private_info == PrivateInfoType::kAllPlayers) {
      auto out = allocator->Get("private_cards", {n, cards});
      for (int p = 0; p < n; ++p) if (v.private_cards[p] >= 0) out.at(p, v.private_cards[p]) = 1;
    }
    if (t_.public_info) {
      auto pub = allocator->Get("community_card", {cards});
      if (v.public_card >= 0) pub.at(v.public_card) = 1;
      if (t_.perfect_recall) {
        auto out = allocator->Get("betting", {2, 3 * n - 2, 2});
        for (int round = 0; round < 2; ++round) {
          const std::vector<int>& seq = round == 0 ? v.round1 : v.round2;
          for (size_t i = 0; i < seq.size(); ++i) {
            if (seq[i] == 1) out.at(round, (int)i, 0) = 1;         // call = 10
            else if (seq[i] == 2) out.at(round, (int)i, 1) = 1;    // raise = 01
          }
        }
      } else {
        auto out = allocator->Get("pot_contribution", {n});
        for (int p = 0; p < n; ++p) out.at(p) = v.ante[p];
      }
    }
  }

  std::string StringFrom(const State& observed_state, int player) const override {
    const B200State& state = open_spiel::down_cast<const B200State&>(observed_state);
    const B200State::PokerView v = state.Poker();
    SPIEL_CHECK_GE(player, 0);
    SPIEL_CHECK_LT(player, v.num_players);
    const bool single = t_.private_info == PrivateInfoType::kSinglePlayer;
    const bool no_private = t_.private_info == PrivateInfoType::kNone;
    const int n = v.num_players;
    const int hist = (int)state.History().size();
    std::string s;
    if (gid_ == B2S_KUHN_POKER) {
      if (single) {
        if (t_.perfect_recall || t_.public_info) {
          if (hist > player) s += std::to_string(v.private_cards[player]);
        } else if (hist == 1 + player) {
          s += "Received card " + std::to_string(v.private_cards[player]);
        }
      }
      if (t_.public_info) {
        if (t_.perfect_recall) {
          for (int a : v.round1) s += a ? 'b' : 'p';
        } else if (no_private) {
          if (hist == 0) s += "start game";
          else if (hist > n) s += v.round1.back() ? "Bet" : "Pass";
        } else if (hist > player) {
          for (int p = 0; p < n; ++p) s += std::to_string(v.ante[p]);
        }
      }
      if (t_.public_info && no_private && hist > 0 && hist <= n) s += "Deal to player " + std::to_string(hist - 1);
      return s;
    }
    auto join = [](const std::vector<int>& x, const char* sep) {
      std::string t;
      for (size_t i = 0; i < x.size(); ++i) { if (i) t += sep; t += std::to_string(x[i]); }
      return t;
    };
    if (single) {
      s += "[Observer: " + std::to_string(player) + "][Private: " + std::to_string(v.private_cards[player]) + "]";
    } else if (t_.private_info == PrivateInfoType::kAllPlayers) {
      s += "[Privates: " + join(v.private_cards, "") + "]";
    }
    if (t_.public_info) {
      s += "[Round " + std::to_string(v.round) + "][Player: " + std::to_string(v.cur_player) + "][Pot: " + std::to_string(v.pot) + "][Money:";
      for (int p = 0; p < n; ++p) s += " " + Num(v.money[p]);
      s += "]";
      if (v.public_card >= 0) s += "[Public: " + std::to_string(v.public_card) + "]";
      if (t_.perfect_recall) s += "[Round1: " + join(v.round1, " ") + "][Round2: " + join(v.round2, " ") + "]";
      else s += "[Ante: " + join(v.ante, " ") + "]";
    }
    return s;
  }

 private:
  int gid_;
  IIGObservationType t_;
};
}  // namespace

std::shared_ptr<Observer> B200Game::MakeObserver(absl::optional<IIGObservationType> iig_obs_type,
                                                 const GameParameters& params) const {
  if ((gid_ == B2S_KUHN_POKER || gid_ == B2S_LEDUC_POKER) && params.empty())
    return std::make_shared<B200PokerObserver>(gid_, iig_obs_type.value_or(kDefaultObsType));
  return Game::MakeObserver(iig_obs_type, params);
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
    case B2S_HAVANNAH: {          // havannah.cc:203-206, Move::ToString :160-164
      const int d = gi.obs_shape[1];
      return std::string(1, (char)('a' + a % d)) + std::to_string(a / d + 1);
    }
    case B2S_Y: {                 // y.cc:143-145, Move::ToString :110-114
      const int n = gi.obs_shape[1];
      return std::string(1, (char)('a' + a % n)) + std::to_string(a / n + 1);
    }
    case B2S_OTHELLO:             // othello.cc:226-234, Move::ToString :120-122
      if (a == 64) return "pass";
      return std::string(1, "abcdefgh"[a % 8]) + std::to_string(1 + a / 8);
    case B2S_MNK: {               // mnk.cc:246-249
      TouchBoardParams();
      const int cols = gi.obs_shape[2];
      return std::string(player == 0 ? "x" : "o") + "(" + std::to_string(a / cols) + "," + std::to_string(a % cols) + ")";
    }
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
  bgame().TouchLineParam();
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
void B200State::FromBatchLane(void* batch, int64_t lane) {
  Check(b2s_state_get(batch, lane, blob_.data(), rules().blob_bytes()));
  undo_.clear();
  if (bgame().gid() == B2S_KUHN_POKER) {      // the packed kuhn state is the action history: rebuild history_ from it
    b2s_host::Decoded d;
    rules().Decode(blob_.data(), &d);
    history_.clear();
    for (int p = 0; p < num_players_ && d.private_card[p] >= 0; ++p) history_.push_back({kChancePlayerId, d.private_card[p]});
    for (size_t k = 0; k < d.round1.size(); ++k) history_.push_back({(Player)(k % num_players_), d.round1[k]});
    move_number_ = (int)history_.size();
  }
}

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
    case B2S_HAVANNAH: {          // havannah.cc:212-279 (ansi_color_output = false)
      const int dd = gi.obs_shape[1], size = (dd + 1) / 2;
      s = std::string(size + 3, ' ');
      for (int x = 0; x < size; ++x) { s += ' '; s += (char)('a' + x); }
      s += '\n';
      for (int y = 0; y < dd; ++y) {
        s += std::string(std::abs(size - 1 - y) + 1 + ((y + 1) < 10), ' ');
        s += std::to_string(y + 1);
        bool found_last = false;
        const int start_x = y < size ? 0 : y - size + 1, end_x = y < size ? size + y : dd;
        for (int x = start_x; x < end_x; ++x) {
          const int xy = x + y * dd;
          if (found_last) { s += ']'; found_last = false; }
          else if (d.last_move == xy) { s += '['; found_last = true; }
          else s += ' ';
          s += ".O@"[d.cells[xy]];
        }
        if (found_last) s += ']';
        if (y < size - 1) { s += ' '; s += (char)('a' + size + y); }
        s += '\n';
      }
      return s;
    }
    case B2S_Y: {                 // y.cc:147-212 (ansi_color_output = false)
      const int n = gi.obs_shape[1];
      s = " ";
      for (int x = 0; x < n; ++x) { s += ' '; s += (char)('a' + x); }
      s += '\n';
      for (int y = 0; y < n; ++y) {
        s += std::string(y + ((y + 1) < 10), ' ');
        s += std::to_string(y + 1);
        bool found_last = false;
        for (int x = 0; x < n - y; ++x) {
          const int xy = x + y * n;
          if (found_last) { s += ']'; found_last = false; }
          else if (d.last_move == xy) { s += '['; found_last = true; }
          else s += ' ';
          s += ".O@"[d.cells[xy]];
        }
        if (found_last) s += ']';
        s += '\n';
      }
      return s;
    }
    case B2S_OTHELLO: {           // othello.cc:245-260
      const std::string cols = "  a b c d e f g h  ";
      s = IsTerminal() ? std::string("Terminal State:\n") : std::string(d.to_play == 0 ? "Black (x)" : "White (o)") + " to play:\n";
      s += cols + "\n";
      for (int r = 0; r < 8; ++r) {
        s += std::to_string(r + 1) + " ";
        for (int col = 0; col < 8; ++col) { s += "-xo"[d.cells[r * 8 + col]]; s += ' '; }
        s += std::to_string(r + 1) + "\n";
      }
      return s + cols;
    }
    case B2S_MNK: {               // mnk.cc:193-205
      const int rows = gi.obs_shape[1], cols = gi.obs_shape[2];
      for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) s += ".xo"[d.cells[r * cols + c]];
        if (r < rows - 1) s += "\n";
      }
      return s;
    }
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
      const PokerView v = Poker();
      const char* more = num_players_ > 2 ? " [...]" : "";
      s = "Round: " + std::to_string(v.round) + "\nPlayer: " + std::to_string(v.cur_player) + "\nPot: " + std::to_string(v.pot) +
          "\nMoney (player_0 player_1" + more + "):";
      for (int p = 0; p < num_players_; ++p) s += " " + Num(v.money[p]);
      s += std::string("\nCards (public player_0 player_1") + more + "): " + std::to_string(v.public_card) + " ";
      for (int p = 0; p < num_players_; ++p) s += std::to_string(v.private_cards[p]) + " ";
      s += "\nRound 1 sequence: ";
      for (size_t i = 0; i < v.round1.size(); ++i) { if (i) s += ", "; s += kNames[v.round1[i]]; }
      s += "\nRound 2 sequence: ";
      for (size_t i = 0; i < v.round2.size(); ++i) { if (i) s += ", "; s += kNames[v.round2[i]]; }
      s += "\n";
      return s;
    }
  }
  return HistoryString();
}

// The poker games answer through their observers (kuhn_poker.cc:285-327, leduc_poker.cc:516-544); the board games
// return HistoryString() / ToString() (e.g. connect_four.cc:287-297).
std::string B200State::InformationStateString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  const int gid = bgame().gid();
  if (gid == B2S_KUHN_POKER || gid == B2S_LEDUC_POKER) return B200PokerObserver(gid, kInfoStateObsType).StringFrom(*this, player);
  return HistoryString();
}

std::string B200State::ObservationString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  const int gid = bgame().gid();
  if (gid == B2S_KUHN_POKER || gid == B2S_LEDUC_POKER) return B200PokerObserver(gid, kDefaultObsType).StringFrom(*this, player);
  return ToString();
}

B200State::PokerView B200State::Poker() const {
  PokerView v;
  v.num_players = num_players_;
  const int gid = bgame().gid();
  if (gid == B2S_KUHN_POKER) {
    v.private_cards.assign(num_players_, -1);
    for (int p = 0; p < num_players_ && p < (int)history_.size(); ++p) v.private_cards[p] = (int)history_[p].action;
    for (int i = num_players_; i < (int)history_.size(); ++i) v.round1.push_back((int)history_[i].action);
    float obs[32];
    rules().Tensor(blob_.data(), 0, 0, obs);                   // pot contributions are the last num_players_ entries
    const int off = rules().info().observation_tensor_size - num_players_;
    for (int p = 0; p < num_players_; ++p) v.ante.push_back((int)obs[off + p]);
    return v;
  }
  SPIEL_CHECK_EQ(gid, (int)B2S_LEDUC_POKER);
  b2s_host::Decoded d;
  rules().Decode(blob_.data(), &d);
  constexpr int kInvalidCard = -10000, kStartingMoney = 100;   // leduc_poker.h:60,68
  v.round = d.round;
  v.cur_player = d.cur_player;
  v.public_card = d.public_card < 0 ? kInvalidCard : d.public_card;
  v.pot = 0;
  for (int p = 0; p < num_players_; ++p) {
    v.private_cards.push_back(d.private_card[p] < 0 ? kInvalidCard : d.private_card[p]);
    v.ante.push_back(d.ante[p]);
    v.pot += d.ante[p];
    // pot_ / money_ (leduc_poker.cc:628-678, 702-706): antes leave the stacks for the pot; ResolveWinner pays the pot out
    v.money.push_back(kStartingMoney - d.ante[p]);
  }
  v.round1 = d.round1; v.round2 = d.round2;
  if (IsTerminal()) {
    float ret[8];
    rules().Returns(blob_.data(), ret);
    for (int p = 0; p < num_players_; ++p) v.money[p] = kStartingMoney + (double)ret[p];      // Returns = money - starting money
    v.pot = 0;
  }
  return v;
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
  if (name == "mnk") return std::shared_ptr<const Game>(new mnk::MNKGame(params));
  if (name == "othello") return std::shared_ptr<const Game>(new othello::OthelloGame(params));
  if (name == "y") return std::shared_ptr<const Game>(new y_game::YGame(params));
  if (name == "havannah") return std::shared_ptr<const Game>(new havannah::HavannahGame(params));
  SpielFatalError("b200: no stock game " + name);
}
}  // namespace

void RegisterB200Games() {
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"tic_tac_toe", "connect_four", "breakthrough", "hex", "go", "kuhn_poker", "leduc_poker", "mnk", "othello", "y", "havannah"}) {
      if (!IsGameRegistered(name)) continue;                 // a build without that stock game
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
