"""Scratch timing of the fused step kernel (not the bench contract; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
import synth
from nav_gym_b200.batched_env import BatchedNavGym, MapPool

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kind = sys.argv[3] if len(sys.argv) > 3 else 'indoor'
    rng = np.random.RandomState(0)
    m = synth.indoor_map(rng) if kind == 'indoor' else synth.outdoor_map(rng)
    pool = MapPool([m], 'cuda:0')
    edt = pool.edt(0).cpu().numpy()
    start = synth.free_poses(rng, m, B, 20, edt)
    goal = synth.free_poses(rng, m, B, 20, edt)
    theta = rng.uniform(0, 2 * np.pi, B)
    env = BatchedNavGym(B, pool, max_disc=P, seed=1)
    env.set_state(start, goal, theta, noise_std=np.full(B, 0.02, np.float32))
    discs = ndisc = None
    if P:
        d = np.zeros((B, P, 3), np.float32)
        d[:, :, :2] = start[:, None, :] + rng.uniform(-8, 8, (B, P, 2))
        d[:, :, 2] = 0.3
        discs = torch.from_numpy(d).cuda(); ndisc = torch.full((B,), P, dtype=torch.int32, device='cuda')
    env.reset(discs, ndisc)
    act = torch.from_numpy(rng.uniform([0, -0.64], [0.5, 0.64], (B, 2)).astype(np.float32)).cuda()
    for _ in range(20):
        env.step(act, discs, ndisc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        env.step(act, discs, ndisc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('B=%d P=%d %s: %.3f ms/step  %.3e env-steps/s  %.3e rays/s  crash frac %.3f' % (
        B, P, kind, ms, B / ms * 1e3, B * 512 / ms * 1e3, env.is_crash.float().mean().item()))

main()
