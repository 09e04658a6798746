"""The synthetic worlds of BASELINE.json's configurations, one definition shared by
bench.py, the parity tests at those sizes and the tools/ scripts.

  C1/C2  load_bench_world(): the reference's own create_indoor_map(3, 100) under
         np.random.seed(0) (map_generator.py:97-123) + a 65 536-tuple spawn pool obeying the
         reference's episode law, from the committed fixture tests/golden/bench_world.npz
         (minted by oracle/make_bench_world.py) -- both arms of bench.py step this world
  C3     c3_world(): 2000 x 2000-cell (100 m) outdoor map, 250 boxes; 20 scripted pedestrians
         per environment (legs + boxes, or trunk discs)
  C4     c4_world(): 8 indoor + 8 outdoor maps of the reference's parameter ranges
         (__init__.py:26-36), environments spread uniformly over them, 5..15 pedestrians, map
         re-drawn at every auto-reset
Reset-time host code; nothing here is on the per-step path.
"""
import os

import numpy as np

from . import _lib, maps

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH_WORLD = os.path.join(_ROOT, 'tests', 'golden', 'bench_world.npz')


def load_bench_world(path=BENCH_WORLD):
    """-> (map_info, pool float64 [65536, 5] = sx, sy, gx, gy, theta).  Pure numpy: usable by the
    CPU reference arm without the CUDA library."""
    z = np.load(path)
    H, W = [int(v) for v in z['map_shape']]
    occ = np.unpackbits(z['map_bits'])[:H * W].reshape(H, W)
    data = np.zeros((H, W), np.int8)
    data[occ > 0] = 100
    m = dict(data=data, origin=(float(z['origin'][0]), float(z['origin'][1])),
             resolution=float(z['resolution']), width=W, height=H)
    if m['origin'] == (0.0, 0.0):
        m['origin'] = (0, 0)
    pool = np.column_stack([(z['pool_cells'].astype(np.float64) + 0.5) * 0.25, z['pool_theta']])
    return m, pool


def scripted_pedestrians(m, n_env, P, rng, bank=512):
    """[n_env, P, NAVGYM_PED_F] rows of waypoint walkers following the reference's spawn law
    (env.py:786-806) minus the 4 m rule (the bank is drawn far from any robot)."""
    rows = maps.spawn_pedestrians(m, (-100.0, -100.0), bank, rng)
    return rows[rng.randint(len(rows), size=(n_env, P))]


def c3_world(device, envs=16384, P=20, seed=3, pool_n=32768):
    """-> (map_info, MapPool with its filtered spawn pool, pedestrian rows [envs, P, 16])."""
    from .batched_env import MapPool, filter_spawn_pool
    rng = np.random.RandomState(seed)
    m = maps.create_large_outdoor_map(rng)
    pool = filter_spawn_pool(m, maps.spawn_pool(m, pool_n, rng), device)
    mp = MapPool([m], device, spawn_pools=[pool])
    return m, mp, scripted_pedestrians(m, envs, P, rng, bank=4096)


def c4_world(device, envs=8192, seed=4, pool_n=8192, P=15, shard=0):
    """-> (maps list, MapPool, map_id [envs], pedestrian rows [envs, P, 16], nped [envs]).  The
    16 maps and their pools are the same on every shard; map ids / pedestrians are drawn per
    shard."""
    from .batched_env import MapPool, filter_spawn_pool
    rng = np.random.RandomState(seed)
    ms = [maps.create_indoor_map(rng.randint(3, 5), rng.randint(80, 151), rng) for _ in range(8)]
    ms += [maps.create_outdoor_map(10, rng.uniform(0.3, 1.0), rng) for _ in range(8)]
    pools = []
    for m in ms:
        lo, hi = (10, 20) if m['width'] > 400 else (5, 15)   # a 20 m outdoor square cannot hold 10-20 m goals at scale
        p = maps.spawn_pool(m, pool_n, rng, min_goal_dist=lo, max_goal_dist=hi)
        pools.append(filter_spawn_pool(m, p, device))
    mp = MapPool(ms, device, spawn_pools=pools)
    rng = np.random.RandomState(seed * 1000 + 17 + shard)
    map_id = rng.randint(0, len(ms), envs).astype(np.int32)
    banks = [maps.spawn_pedestrians(m, (-100.0, -100.0), 256, rng) for m in ms]
    peds = np.zeros((envs, P, _lib.PED_F), np.float32)
    for i in range(len(ms)):
        sel = np.where(map_id == i)[0]
        peds[sel] = banks[i][rng.randint(len(banks[i]), size=(len(sel), P))]
    nped = rng.randint(5, 16, envs).astype(np.int32)
    return ms, mp, map_id, peds, nped
