import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import open_spiel_b200 as b2
from oracle_lib import OracleGame
for gs in ["hex(board_size=8)", "hex(board_size=5,plain_obs_tensor=True)", "hex(board_size=2)", "hex(board_size=3)"]:
    g = b2.load_game(gs)
    o = OracleGame(gs).new_initial_state()
    e = o.observation_tensor(0)
    for nl in (1, 2, 3, 4, 32, 64):
        b = g.new_batch(nl)
        out = torch.full((nl, g.observation_tensor_size()), 7.0, dtype=torch.float32, device="cuda")
        d = b.observation_tensor(0, out=out).cpu().numpy()
        for lane in sorted(set([0, nl - 1])):
            bad = np.nonzero(d[lane] != e)[0]
            print(gs, "n", nl, "lane", lane, "nbad", len(bad), bad[:12].tolist(), d[lane][bad[:6]].tolist())
