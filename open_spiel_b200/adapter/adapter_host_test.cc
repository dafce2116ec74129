// CPU-only test of the drop-in boundary: all seven games obtained through open_spiel::LoadGame after
// RegisterB200Games(), (1) run through the REFERENCE'S OWN harness (open_spiel/tests/basic_tests.cc: RandomSimTest,
// RandomSimTestWithUndo, CheckChanceOutcomes, RandomSimTestCustomObserver — the calls the stock *_test.cc files make,
// e.g. breakthrough_test.cc:52-53, go_test.cc:39-40, kuhn_poker_test.cc:31-40, leduc_poker_test.cc:35-48,
// hex_test.cc:73-78), and (2) played in lock-step against the stock C++ game on every observable, strings included.
// Scalar State methods run on the host build of the rule cores (host_rules.h), so this needs no GPU.
#include <iostream>
#include <random>

#include "b200_games.h"
#include "open_spiel/spiel.h"
#include "open_spiel/tests/basic_tests.h"

using namespace open_spiel;

static int LockStep(const Game& ours, const Game& stock, int games, std::mt19937* rng) {
  int steps = 0;
  const GameType& type = stock.GetType();
  SPIEL_CHECK_EQ(ours.ToString(), stock.ToString());
  SPIEL_CHECK_EQ(ours.NumDistinctActions(), stock.NumDistinctActions());
  SPIEL_CHECK_EQ(ours.NumPlayers(), stock.NumPlayers());
  SPIEL_CHECK_EQ(ours.MaxGameLength(), stock.MaxGameLength());
  SPIEL_CHECK_EQ(ours.MaxChanceOutcomes(), stock.MaxChanceOutcomes());
  SPIEL_CHECK_EQ(ours.MinUtility(), stock.MinUtility());
  SPIEL_CHECK_EQ(ours.MaxUtility(), stock.MaxUtility());
  SPIEL_CHECK_TRUE(ours.ObservationTensorShape() == stock.ObservationTensorShape());
  if (type.provides_information_state_tensor)
    SPIEL_CHECK_TRUE(ours.InformationStateTensorShape() == stock.InformationStateTensorShape());
  SPIEL_CHECK_TRUE(ours.GetParameters() == stock.GetParameters());
  for (int g = 0; g < games; ++g) {
    auto a = ours.NewInitialState();
    auto b = stock.NewInitialState();
    while (true) {
      SPIEL_CHECK_EQ(a->IsTerminal(), b->IsTerminal());
      SPIEL_CHECK_EQ(a->CurrentPlayer(), b->CurrentPlayer());
      SPIEL_CHECK_TRUE(a->LegalActions() == b->LegalActions());
      std::vector<double> ra = a->Returns(), rb = b->Returns();
      SPIEL_CHECK_EQ(ra.size(), rb.size());
      for (size_t i = 0; i < ra.size(); ++i) {
        SPIEL_CHECK_EQ(ra[i], rb[i]);
        SPIEL_CHECK_EQ(std::signbit(ra[i]), std::signbit(rb[i]));      // hex returns {0, -0} before the end
      }
      SPIEL_CHECK_TRUE(a->Rewards() == b->Rewards());
      SPIEL_CHECK_EQ(a->ToString(), b->ToString());
      SPIEL_CHECK_EQ(a->Serialize(), b->Serialize());
      for (Player p = 0; p < stock.NumPlayers(); ++p) {
        if (type.provides_observation_tensor) SPIEL_CHECK_TRUE(a->ObservationTensor(p) == b->ObservationTensor(p));
        if (type.provides_observation_string) SPIEL_CHECK_EQ(a->ObservationString(p), b->ObservationString(p));
        if (type.provides_information_state_string) SPIEL_CHECK_EQ(a->InformationStateString(p), b->InformationStateString(p));
        if (type.provides_information_state_tensor) SPIEL_CHECK_TRUE(a->InformationStateTensor(p) == b->InformationStateTensor(p));
        SPIEL_CHECK_TRUE(a->LegalActions(p) == b->LegalActions(p));
      }
      if (a->IsTerminal()) break;
      if (a->IsChanceNode()) SPIEL_CHECK_TRUE(a->ChanceOutcomes() == b->ChanceOutcomes());
      auto la = b->LegalActions();
      for (Action act : la) SPIEL_CHECK_EQ(a->ActionToString(act), b->ActionToString(act));
      Action act = la[(*rng)() % la.size()];
      a->ApplyAction(act);
      b->ApplyAction(act);
      ++steps;
    }
    SPIEL_CHECK_TRUE(a->History() == b->History());
  }
  // and again after play: the stock mnk records "k" on the first DoApplyAction (mnk.h:120-123), and so does the drop-in
  SPIEL_CHECK_TRUE(ours.GetParameters() == stock.GetParameters());
  SPIEL_CHECK_EQ(ours.ToString(), stock.ToString());
  return steps;
}

int main() {
  // (game string, lock-step games, RandomSimTest sims, RandomSimTestWithUndo sims)
  struct Case { const char* game; int lockstep, sims, undo; };
  const Case cases[] = {
      {"tic_tac_toe", 60, 100, 2},
      {"connect_four", 40, 100, 2},
      {"connect_four(rows=4,columns=5,x_in_row=3)", 20, 10, 1},                  // connect_four_test.cc:325
      {"connect_four(rows=7,columns=8,x_in_row=5)", 20, 10, 1},                  // connect_four_test.cc:377
      {"connect_four(egocentric_obs_tensor=true)", 10, 5, 0},
      {"breakthrough", 20, 100, 1},                                              // breakthrough_test.cc:52-53
      {"breakthrough(rows=6,columns=6)", 10, 10, 1},
      {"breakthrough(rows=5,columns=4)", 10, 10, 1},
      {"hex(num_cols=5,num_rows=5)", 20, 100, 1},                                // hex_test.cc:73-78
      {"hex", 5, 5, 1},
      {"hex(num_cols=2,num_rows=3)", 10, 10, 0},
      {"hex(num_cols=2,num_rows=2)", 10, 10, 0},
      {"hex(swap=true)", 10, 10, 1},
      {"hex(plain_obs_tensor=true,swap=true)", 10, 10, 0},
      {"hex(board_size=4,string_rep=explicit)", 10, 5, 0},
      {"go(board_size=9,komi=7.5)", 6, 3, 3},                                    // go_test.cc:36-40 uses komi 7.5, size 19 -> 9 here
      {"go(board_size=7,komi=4.5)", 6, 3, 1},
      {"go(board_size=5)", 10, 5, 1},
      {"go(board_size=3,max_game_length=30)", 20, 10, 1},
      {"go(board_size=2)", 20, 10, 1},
      {"havannah", 20, 10, 1},                                                   // havannah_test.cc: RandomSimTest on sizes 3..8, with / without swap
      {"havannah(board_size=4)", 100, 100, 1},
      {"havannah(board_size=4,swap=true)", 100, 100, 1},
      {"havannah(board_size=3,swap=true)", 60, 60, 1},
      {"y(board_size=9)", 100, 60, 1},                                           // y_test.cc: RandomSimTest on sizes up to 11
      {"y(board_size=11)", 30, 30, 1},
      {"y(board_size=3)", 60, 60, 1},
      {"othello", 100, 60, 1},                                                   // othello_test.cc:30-34
      {"mnk", 6, 5, 1},                                                          // mnk_test.cc: RandomSimTest
      {"mnk(m=3,n=3,k=3)", 60, 50, 1},
      {"mnk(m=7,n=5,k=4)", 20, 20, 1},
      {"mnk(m=15,n=2,k=6)", 10, 10, 0},
      {"kuhn_poker", 60, 100, 1},                                                // kuhn_poker_test.cc:31-32
      {"kuhn_poker(players=3)", 60, 50, 1},                                      // kuhn_poker_test.cc:33-38 (2..4 players)
      {"kuhn_poker(players=4)", 40, 50, 1},
      {"kuhn_poker(players=5)", 20, 20, 1},
      {"leduc_poker", 60, 100, 1},                                               // leduc_poker_test.cc:35
      {"leduc_poker(starting_player=1)", 30, 20, 1},
      {"leduc_poker(players=3)", 40, 50, 1},                                     // leduc_poker_test.cc:36-46
      {"leduc_poker(players=3,starting_player=2)", 20, 10, 1},
      {"leduc_poker(players=4)", 20, 20, 1},
  };
  std::vector<std::shared_ptr<const Game>> stock;
  for (const Case& c : cases) stock.push_back(LoadGame(c.game));     // built by the stock factories: names not yet taken over
  std::shared_ptr<const Game> stock_go19 = LoadGame("go(board_size=19)");
  b200::RegisterB200Games();
  b200::RegisterB200Games();                                           // idempotent
  std::mt19937 rng(7);
  long total = 0;
  for (size_t i = 0; i < stock.size(); ++i) {
    const Case& c = cases[i];
    std::shared_ptr<const Game> ours = LoadGame(c.game);
    SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(ours.get()) != nullptr);     // LoadGame now returns the adapter
    SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(stock[i].get()) == nullptr);
    total += LockStep(*ours, *stock[i], c.lockstep, &rng);
    testing::RandomSimTest(*ours, c.sims);                             // the reference's own harness on the drop-in
    if (c.undo) testing::RandomSimTestWithUndo(*ours, c.undo);
    if (ours->GetType().chance_mode != GameType::ChanceMode::kDeterministic) {
      testing::CheckChanceOutcomes(*ours);                             // kuhn_poker_test.cc:74
      testing::RandomSimTestCustomObserver(*ours, ours->MakeObserver(kDefaultObsType, {}));   // kuhn_poker_test.cc:40
      testing::RandomSimTestCustomObserver(*ours, ours->MakeObserver(kInfoStateObsType, {}));
    }
    std::cout << "ok " << c.game << std::endl;
  }
  // parameter sets the packed layouts cannot hold are served by the stock game (the previous factory)
  for (const char* g : {"go(board_size=19)", "kuhn_poker(players=6)", "leduc_poker(players=5)", "leduc_poker(suit_isomorphism=true)",
                        "hex(board_size=13)", "connect_four(rows=12,columns=12)", "y", "y(board_size=9,ansi_color_output=true)", "havannah(board_size=9)", "havannah(board_size=4,ansi_color_output=true)",
                        "mnk(m=16,n=4,k=3)"}) {
    std::shared_ptr<const Game> fb = LoadGame(g);
    SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(fb.get()) == nullptr);
    testing::RandomSimTest(*fb, 1);
  }
  SPIEL_CHECK_EQ(LoadGame("go(board_size=19)")->ToString(), stock_go19->ToString());
  // known answers of the stock tests through the drop-in: connect_four_test.cc FastLoss / draw positions need the
  // concrete ConnectFourState class; the generic ones are
  {
    std::shared_ptr<const Game> game = LoadGame("kuhn_poker");
    auto s = game->NewInitialState();
    s->ApplyAction(2); s->ApplyAction(1); s->ApplyAction(1); s->ApplyAction(1);        // deal 2,1; bet, bet
    SPIEL_CHECK_TRUE(s->IsTerminal());
    SPIEL_CHECK_TRUE(s->Returns() == (std::vector<double>{2, -2}));
    std::shared_ptr<const Game> go = LoadGame("go(board_size=9)");
    SPIEL_CHECK_EQ(go->NumDistinctActions(), 82);                                      // go_test.cc: board_size^2 + 1
    auto t = go->DeserializeState(go->NewInitialState()->Serialize());
    SPIEL_CHECK_EQ(t->ToString(), go->NewInitialState()->ToString());
  }
  std::cout << "adapter_host_test ok: " << total << " lock-step transitions against the stock games" << std::endl;
  return 0;
}
