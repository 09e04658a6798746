"""Throughput of the other BASELINE.json configurations (bench.py measures configs[1]):
  c3  16384 envs, 2000x2000-cell outdoor map, 20 device-scripted pedestrians per env
  c4  8192 envs per GPU (65536 / 8), 8 indoor + 8 outdoor maps, 5..15 pedestrians, map re-drawn
      at auto-reset
  c5  32768 envs on the c2 world stepped by a torch MLP policy on the device (rollout loop)
Device-resident numbers (CUDA events, no L2 flush), one JSON line per config."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200 import maps, _lib
from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool


def timed(env, policy, steps, warmup):
    obs = env.obs
    for _ in range(warmup):
        env.step(policy(obs))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        env.step(policy(obs))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def random_policy(B, dev, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    bank = torch.rand(32, B, 2, device=dev, generator=g) * torch.tensor([0.5, 1.28], device=dev) + torch.tensor([0, -0.64], device=dev)
    it = [0]
    def f(obs):
        it[0] += 1
        return bank[it[0] % 32]
    return f


def peds_for(m, rows, P, rng, nped=None):
    out = np.zeros((len(rows), P, _lib.PED_F), np.float32)
    bank = maps.spawn_pedestrians(m, (-100.0, -100.0), 4096, rng)   # far from any robot: law minus the 4 m rule
    for e in range(len(rows)):
        out[e] = bank[rng.randint(len(bank), size=P)]
    return out


def c3(dev, steps, warmup, B=16384, P=20):
    rng = np.random.RandomState(3)
    m = maps.create_large_outdoor_map(rng)
    pool = filter_spawn_pool(m, maps.spawn_pool(m, 32768, rng), dev)
    mp = MapPool([m], dev, spawn_pools=[pool])
    res = {}
    for mode, trunk in (('legs+boxes', False), ('trunk discs', True)):
        env = BatchedNavGym(B, mp, device=dev, seed=5, auto_reset=True)
        env.reset_from_spawn_pool(np.random.RandomState(1))
        rows = env.state[:2].T.cpu().numpy()
        env.attach_pedestrians(peds_for(m, rows, P, rng), trunk_mode=trunk)
        env.reset()
        ms = timed(env, random_policy(B, dev), steps, warmup)
        res[mode] = dict(ms_per_step=ms, env_steps_per_s=B / ms * 1e3, rays_per_s=B * 512 / ms * 1e3)
    return dict(config='c3', envs=B, map='outdoor 2000x2000 cells, 250 boxes', pedestrians=P, results=res)


def c4(dev, steps, warmup, B=8192):
    rng = np.random.RandomState(4)
    ms_, pools = [], []
    for i in range(8):
        ms_.append(maps.create_indoor_map(rng.randint(3, 5), rng.randint(80, 151), rng))
    for i in range(8):
        ms_.append(maps.create_outdoor_map(10, rng.uniform(0.3, 1.0), rng))
    for m in ms_:
        lo = (10, 20) if m['width'] > 400 else (5, 15)
        p = maps.spawn_pool(m, 8192, rng, min_goal_dist=lo[0], max_goal_dist=lo[1])
        pools.append(filter_spawn_pool(m, p, dev))
    mp = MapPool(ms_, dev, spawn_pools=pools)
    map_id = rng.randint(0, 16, B).astype(np.int32)
    env = BatchedNavGym(B, mp, device=dev, map_id=map_id, seed=6, auto_reset=True, resample_map=True)
    env.reset_from_spawn_pool(np.random.RandomState(2))
    P = 15
    peds = np.zeros((B, P, _lib.PED_F), np.float32)
    banks = [maps.spawn_pedestrians(m, (-100.0, -100.0), 512, rng) for m in ms_]
    for e in range(B):
        peds[e] = banks[map_id[e]][rng.randint(512, size=P)]
    nped = rng.randint(5, 16, B).astype(np.int32)
    env.attach_pedestrians(peds, nped=nped)
    env.reset()
    ms = timed(env, random_policy(B, dev), steps, warmup)
    return dict(config='c4 (one GPU share)', envs=B, maps='8 indoor + 8 outdoor', pedestrians='5..15',
                ms_per_step=ms, env_steps_per_s=B / ms * 1e3, rays_per_s=B * 512 / ms * 1e3)


def c5(dev, steps, warmup, B=32768):
    from bench import build_world
    m, pool = build_world(0)
    pool = filter_spawn_pool(m, pool, dev)
    mp = MapPool([m], dev, spawn_pools=[pool])
    env = BatchedNavGym(B, mp, device=dev, seed=7, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(3))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(519, 256), torch.nn.Tanh(), torch.nn.Linear(256, 256), torch.nn.Tanh(),
                              torch.nn.Linear(256, 2)).to(dev).to(torch.bfloat16)
    lo = torch.tensor([0.0, -0.64], device=dev); hi = torch.tensor([0.5, 0.64], device=dev)
    @torch.no_grad()
    def policy(obs):
        x = obs.to(torch.bfloat16)
        x[:, :512] = x[:, :512] / 25.0
        a = torch.sigmoid(net(x).float())
        return lo + (hi - lo) * a
    ms = timed(env, policy, steps, warmup)
    ms_env = timed(env, random_policy(B, dev), steps, warmup)
    return dict(config='c5', envs=B, policy='MLP 519-256-256-2 bf16, obs consumed in place on device',
                ms_per_step=ms, env_steps_per_s=B / ms * 1e3, ms_per_step_env_only=ms_env)


def crowd(dev, steps, warmup, B=4096, P=10):
    """C2's world with the reference's policy-driven pedestrians (PedestrianSim): P pedestrians per
    env, each with its own 512-beam scan and a CNN policy forward per step (random-init weights)."""
    from bench import build_world
    from nav_gym_b200.pedestrians import PedestrianSim
    m, pool = build_world(0)
    pool = filter_spawn_pool(m, pool, dev)
    mp = MapPool([m], dev, spawn_pools=[pool])
    env = BatchedNavGym(B, mp, device=dev, seed=8, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(4))
    precision = os.environ.get('NAVGYM_CROWD_PRECISION', 'tf32')
    sim = PedestrianSim(env, P, seed=8, precision=precision)
    pol = random_policy(B, dev)

    def phase(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for _ in range(warmup):
        sim.step(pol(None))
    ms = phase(lambda: sim.step(pol(None)), steps)
    parts = dict(act=phase(sim.act, 20), robot_step=phase(lambda: env.step(pol(None)), 20), observe=phase(sim.observe, 20))
    return dict(config='crowd', envs=B, pedestrians=P, policy_precision=precision, ms_per_step=ms, env_steps_per_s=B / ms * 1e3,
                pedestrian_steps_per_s=B * P / ms * 1e3, pedestrian_rays_per_s=B * P * 512 / ms * 1e3,
                ms_parts=parts)


if __name__ == '__main__':
    dev = torch.device('cuda:0')
    which = sys.argv[1:] or ['c3', 'c4', 'c5']
    for w in which:
        print(json.dumps(globals()[w](dev, 100, 20)))
