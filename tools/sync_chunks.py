"""Synchronous host-buffer step (navgym_step_batch_host) vs the number of chunks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool
from bench import build_world
B = 4096
m, pool = build_world(0, 65536)
pool = filter_spawn_pool(m, pool, 'cuda:0')
mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
g = torch.Generator(device='cuda'); g.manual_seed(0)
bank = (torch.rand(16, B, 2, device='cuda', generator=g) * torch.tensor([0.5, 1.28], device='cuda') + torch.tensor([0, -0.64], device='cuda')).cpu().pin_memory()
obs_h = torch.empty(B, 519).pin_memory(); rew_h = torch.empty(B).pin_memory(); done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
for chunks in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]:
    env = BatchedNavGym(B, mp, seed=1, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(1))
    for i in range(40): env.step_host(bank[i % 16], obs_h, rew_h, done_h, chunks=chunks)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 400
    for i in range(n): env.step_host(bank[i % 16], obs_h, rew_h, done_h, chunks=chunks)
    dt = time.perf_counter() - t0
    print('chunks %d: %.1f us/step  %.2f M env-steps/s' % (chunks, dt / n * 1e6, B * n / dt / 1e6))
