"""A few steps of BASELINE config C3 / C4 for profiling (ncu launch lists):
    python tools/c3_prof.py [c3|c3trunk|c4] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200 import worlds
from nav_gym_b200.batched_env import BatchedNavGym
from bench import _bank

which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device('cuda:0')
if which.startswith('c3'):
    B, P = 16384, 20
    m, mp, peds = worlds.c3_world(dev, B, P)
    env = BatchedNavGym(B, mp, device=dev, seed=5, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(1))
    env.attach_pedestrians(peds, trunk_mode=which == 'c3trunk')
else:
    B = 8192
    ms_maps, mp, map_id, peds, nped = worlds.c4_world(dev, B)
    env = BatchedNavGym(B, mp, device=dev, map_id=map_id, seed=6, auto_reset=True, resample_map=True)
    env.reset_from_spawn_pool(np.random.RandomState(2))
    env.attach_pedestrians(peds, nped=nped)
env.reset()
bank = _bank(torch, dev, 32, B, 33)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(20):
    env.step(bank[i % 32])
e0.record()
for i in range(steps):
    env.step(bank[i % 32])
e1.record()
torch.cuda.synchronize()
print('%s: %.4f ms/step' % (which, e0.elapsed_time(e1) / steps))
