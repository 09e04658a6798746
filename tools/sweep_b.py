"""ms/step vs number of environments (bench world, auto-reset, Philox noise, no L2 flush)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool
from bench import build_world
m, pool = build_world(0, 65536)
pool = filter_spawn_pool(m, pool, 'cuda:0')
mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
for B in [int(x) for x in sys.argv[1:]] or [1184, 2368, 3552, 4096, 4736, 8192, 16384, 32768]:
    env = BatchedNavGym(B, mp, seed=1, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(1))
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    bank = torch.rand(16, B, 2, device='cuda', generator=g) * torch.tensor([0.5, 1.28], device='cuda') + torch.tensor([0, -0.64], device='cuda')
    for i in range(20): env.step(bank[i % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for i in range(n): env.step(bank[i % 16])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('B=%6d  %.4f ms/step  %.2f M env-steps/s  %.1f ns/env   obs checksum %.6f' % (B, ms, B / ms / 1e3, ms * 1e6 / B, float(env.obs.double().sum())))
