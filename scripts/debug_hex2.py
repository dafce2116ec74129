import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import open_spiel_b200 as b2
g = b2.load_game("hex(board_size=3)")
b = g.new_batch(2)
out = torch.full((2, g.observation_tensor_size()), 7.0, dtype=torch.float32, device="cuda")
d = b.observation_tensor(0, out=out)
torch.cuda.synchronize()
print(d.cpu().numpy()[0][:40])
