"""Run selected entries of bench_configs (cfg2, cfg3, cfg4, cfg5_shard, normal_fill) and print them as JSON.
usage: python tools/run_cfg.py cfg2 [cfg3 ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import htnorm_b200 as h
import bench_configs as bc
which = tuple(sys.argv[1:]) or ("cfg2",)
peak = h.fp64_tensor_peak_tflops(0, 50.0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
res = bc.run_all(h, torch, "cuda:0", peak, float(peaks.get("hbm_gbs", bc.HBM_FALLBACK_GBS)), which)
for k, v in res.items():
    print(k, json.dumps(v))
