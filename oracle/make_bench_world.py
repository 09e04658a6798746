"""Mints tests/golden/bench_world.npz: the world both arms of bench.py step.

TEST / BENCH INFRASTRUCTURE.  Runs in the build container only (needs /root/reference):
    python oracle/make_bench_world.py

  map   the reference's own map_generator.create_indoor_map(corridor_width=3, iterations=100)
        under np.random.seed(0), executed unmodified (map_generator.py:97-123)
  pool  65 536 (start, goal, heading) tuples obeying the reference's episode law (both ends on
        free cells of the 0.25 m cost map env.py:309-332, 10 m < |goal - start| < 20 m :379,
        connected by a path <= 2 x the straight line :761, heading U[0, 2 pi) :763, first scan
        free of discomfort :779-783), drawn with nav_gym_b200.maps.spawn_pool with the BFS
        routed through the CPU checker and filtered with the CPU checker's scan -- the CUDA
        library is not involved.  Stored as cost-map cells (int16) + headings (float64).
Also checks that the product's generator (nav_gym_b200.maps) reproduces the reference's maps bit
for bit for a few seeds and stores those maps' checksums as known answers.
"""
import hashlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/nav_gym/src/nav_gym_env')

import map_generator as ref_maps  # noqa: E402  (the reference's module, unmodified)
from nav_gym_b200 import maps  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'bench_world.npz')
POOL_N = 65536


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    np.random.seed(0)
    m = ref_maps.create_indoor_map(3, 100)
    mine = maps.create_indoor_map(3, 100, np.random.RandomState(0))
    assert np.array_equal(m['data'], mine['data']), 'product generator != reference generator'
    known = {}
    for seed, (kind, args) in enumerate([('indoor', (3, 100)), ('indoor', (4, 150)), ('indoor', (3, 80)),
                                         ('outdoor', (10, 0.7)), ('outdoor', (10, 0.35))]):
        np.random.seed(seed)
        ref = getattr(ref_maps, 'create_%s_map' % kind)(*args)
        got = getattr(maps, 'create_%s_map' % kind)(*args, rng=np.random.RandomState(seed))
        assert np.array_equal(ref['data'], got['data']) and ref['data'].dtype == got['data'].dtype
        known['%s_%s_%s_seed%d' % (kind, args[0], args[1], seed)] = sha(ref['data'])

    maps._BFS_OVERRIDE = orc.grid_bfs          # no CUDA library in this process
    orc.use_all_cores()
    rng = np.random.RandomState(1)
    pool = maps.spawn_pool(m, int(POOL_N * 1.25), rng)
    # discomfort-free first scan (env.py:779-783), by the CPU checker
    keep = []
    for s in range(0, len(pool), 8192):
        p = pool[s:s + 8192]
        o = orc.OracleBatch([m], np.zeros(len(p), np.int32), p[:, 0:2], p[:, 2:4], p[:, 4],
                            params=dict(t_stop=502.0))
        obs = o.reset_obs(want_hits=False)
        keep.append(~(obs[:, :512] < o.dthr[None, :]).any(axis=1))
    pool = pool[np.concatenate(keep)][:POOL_N]
    assert len(pool) == POOL_N, len(pool)
    cells = np.round(pool[:, :4] / 0.25 - 0.5).astype(np.int16)   # (c + 0.5) * 0.25 -> c
    assert np.array_equal((cells + 0.5) * 0.25, pool[:, :4])
    np.savez_compressed(OUT, map_bits=np.packbits(m['data'] > 0), map_shape=np.array(m['data'].shape),
                        resolution=m['resolution'], origin=np.array(m['origin'], np.float64),
                        pool_cells=cells, pool_theta=pool[:, 4],
                        known_names=np.array(sorted(known)), known_sha=np.array([known[k] for k in sorted(known)]))
    print('wrote', OUT, os.path.getsize(OUT), 'bytes; occupancy %.4f; pool %d' % ((m['data'] > 0).mean(), len(pool)))


if __name__ == '__main__':
    main()
