"""HBM roofline of the HER batch kernel (compute_rewards / compute_terminals / compute_info on
stored observations): N rows of 519 float32 + a goal in, 11 bytes out per row.
Algorithmic bytes per row: 2076 (row) + 8 (goal) + 4 + 1 + 1 + 1 + 4 (reward, done, is_success,
is_crash, distance) = 2095."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200 import maps
from nav_gym_b200.batched_env import BatchedNavGym, MapPool

dev = 'cuda:0'
m = maps.create_outdoor_map(10, 0.5, np.random.RandomState(0))
env = BatchedNavGym(8, MapPool([m], dev), device=dev)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6546.9) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6546.9
for N in (1 << 17, 1 << 20):
    g = torch.Generator(device=dev); g.manual_seed(0)
    obs = torch.rand(N, 519, device=dev, generator=g) * 10.0
    goals = torch.rand(N, 2, device=dev, generator=g) * 50.0
    for _ in range(3):
        env.compute_rewards(obs, goals)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.compute_rewards(obs, goals); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = N * 2095 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "her_kernel", "rows": N, "ms": ms, "rows_per_s": N / ms * 1e3, "achieved_GBps": gbs,
                      "peak_GBps": peak, "frac": gbs / peak, "note": "timed through BatchedNavGym.compute_rewards (five output allocations + launch), L2 flushed"}))
