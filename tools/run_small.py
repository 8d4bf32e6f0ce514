"""cfg2-shaped run for profiling: P problems k = 64, k2 = 4 (default 2^18), a few launches"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htnorm_b200 as h
import bench_configs as bc
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
r = bc.cfg2(h, torch, "cuda:0", 6535.1, P)
print(r["ms"], r["samples_per_s"], r["roofline"]["frac"])
