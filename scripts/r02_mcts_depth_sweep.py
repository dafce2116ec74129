#!/usr/bin/env python3
"""Round-2 experiment: go 9x9 MCTS throughput vs number of concurrent trees at depth (>= 4k sims/tree).
Usage: r02_mcts_depth_sweep.py out.jsonl"""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_mcts import run
out = open(sys.argv[1], "a")
run("go(board_size=9)", 256, 16)
for trees, sims in [(2048, 4000), (8192, 4000), (16384, 4000), (32768, 4000), (65536, 2000), (16384, 10000)]:
    r = run("go(board_size=9)", trees, sims)
    print(json.dumps(r), flush=True)
    out.write(json.dumps(r) + "\n"); out.flush()
