// Runs the REFERENCE'S OWN test harness (open_spiel/tests/basic_tests.cc: RandomSimTest — legal actions sorted,
// clone equality, serialization round trip, tensor sizes, returns bounds, game length, ...) on the B200 adapter
// games obtained through open_spiel::LoadGame, then plays the stock C++ game and the adapter in lock-step.
#include <iostream>
#include <random>

#include "b200_games.h"
#include "open_spiel/games/connect_four/connect_four.h"
#include "open_spiel/games/tic_tac_toe/tic_tac_toe.h"
#include "open_spiel/spiel.h"
#include "open_spiel/tests/basic_tests.h"

using namespace open_spiel;

static void LockStep(const Game& ours, const Game& stock, int games, std::mt19937* rng) {
  for (int g = 0; g < games; ++g) {
    auto a = ours.NewInitialState();
    auto b = stock.NewInitialState();
    while (true) {
      SPIEL_CHECK_EQ(a->IsTerminal(), b->IsTerminal());
      SPIEL_CHECK_EQ(a->CurrentPlayer(), b->CurrentPlayer());
      SPIEL_CHECK_TRUE(a->LegalActions() == b->LegalActions());
      SPIEL_CHECK_TRUE(a->Returns() == b->Returns());
      SPIEL_CHECK_EQ(a->ToString(), b->ToString());
      for (Player p = 0; p < 2; ++p) SPIEL_CHECK_TRUE(a->ObservationTensor(p) == b->ObservationTensor(p));
      if (a->IsTerminal()) break;
      auto la = b->LegalActions();
      Action act = la[(*rng)() % la.size()];
      SPIEL_CHECK_EQ(a->ActionToString(act), b->ActionToString(act));
      a->ApplyAction(act);
      b->ApplyAction(act);
    }
    SPIEL_CHECK_TRUE(a->History() == b->History());
  }
}

int main() {
  // stock game objects, built directly from their classes before the names are taken over
  std::shared_ptr<const Game> stock_c4 = LoadGame("connect_four");
  std::shared_ptr<const Game> stock_ttt = LoadGame("tic_tac_toe");
  std::shared_ptr<const Game> stock_c4_small = LoadGame("connect_four(rows=4,columns=5,x_in_row=3)");
  b200::RegisterB200Games();
  std::shared_ptr<const Game> c4 = LoadGame("connect_four");
  std::shared_ptr<const Game> ttt = LoadGame("tic_tac_toe");
  std::shared_ptr<const Game> c4_small = LoadGame("connect_four(rows=4,columns=5,x_in_row=3)");
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(c4.get()) != nullptr);     // LoadGame now returns the adapter
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(ttt.get()) != nullptr);
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(stock_c4.get()) == nullptr);
  SPIEL_CHECK_EQ(c4->NumDistinctActions(), 7);
  SPIEL_CHECK_EQ(c4->MaxGameLength(), 42);
  std::mt19937 rng(7);
  LockStep(*c4, *stock_c4, 40, &rng);
  LockStep(*ttt, *stock_ttt, 40, &rng);
  LockStep(*c4_small, *stock_c4_small, 20, &rng);
  testing::RandomSimTest(*c4, 15);        // the reference's own harness on the drop-in
  testing::RandomSimTest(*ttt, 15);
  testing::RandomSimTest(*c4_small, 10);
  std::cout << "adapter_test ok" << std::endl;
  return 0;
}
