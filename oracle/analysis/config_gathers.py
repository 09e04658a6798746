"""One-off analysis (uses the CPU checker, so it lives under oracle/): EDT gathers per env-step of
the bench configurations -- the constants GATHERS_* of bench.py's roofline_gather blocks.

gathers per scan = 1 (the t = 0 sample, shared by the 512 beams of a scan) + sum over beams of the
march samples after it, with the kernel's 25 m + 2 cells cut-off; poses drawn from each world's
spawn pool (where episodes start), x 1.01 scans per step (crash re-scans + auto-reset first scans,
~1 % of environments per step on the bench world).  CPU only: the spawn-pool BFS is routed through
the checker, and the pools are not filtered for discomfort-free starts (a second-order effect).
    python oracle/analysis/config_gathers.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from nav_gym_b200 import maps, worlds  # noqa: E402
from oracle import oracle as orc  # noqa: E402

maps._BFS_OVERRIDE = orc.grid_bfs
orc.use_all_cores()


def gathers(m, pool, n=1500, seed=0):
    rng = np.random.RandomState(seed)
    rows = pool[rng.randint(len(pool), size=n)]
    dist = orc.edt(np.asarray(m['data']) >= 0.1)
    res = m['resolution']
    ox = np.repeat((rows[:, 0] / res).astype(np.float32).astype(np.int32).astype(np.float32), 512)
    oy = np.repeat((rows[:, 1] / res).astype(np.float32).astype(np.int32).astype(np.float32), 512)
    heads = (np.tile(orc.beam_table(), n) + np.repeat(rows[:, 4].astype(np.float32).astype(np.float64), 512)).astype(np.float32)
    ins = np.column_stack([ox, oy, heads]).astype(np.float32)
    _, steps = orc.calc_range_many(dist, ins, float(m['width'] * m['height']), t_stop=25.0 / res + 2.0, want_steps=True)
    s = steps.reshape(n, 512)
    per_scan = 1 + np.maximum(s - 1, 0).sum(1)
    return float(per_scan.mean()), float(s.mean()), int(s.max())


if __name__ == '__main__':
    m, pool = worlds.load_bench_world()
    g, spr, mx = gathers(m, pool)
    print('C2 bench world: %.0f gathers/scan (%.2f samples/ray, max %d) -> x1.01 = %.0f per env-step' % (g, spr, mx, 1.01 * g))
    rng = np.random.RandomState(3)
    m3 = maps.create_large_outdoor_map(rng)
    g, spr, mx = gathers(m3, maps.spawn_pool(m3, 4096, rng))
    print('C3 outdoor 2000^2: %.0f gathers/scan (%.2f samples/ray, max %d) -> x1.01 = %.0f per env-step' % (g, spr, mx, 1.01 * g))
    rng = np.random.RandomState(4)
    ms = [maps.create_indoor_map(rng.randint(3, 5), rng.randint(80, 151), rng) for _ in range(8)]
    ms += [maps.create_outdoor_map(10, rng.uniform(0.3, 1.0), rng) for _ in range(8)]
    tot = []
    for m4 in ms:
        lo, hi = (10, 20) if m4['width'] > 400 else (5, 15)
        g, spr, mx = gathers(m4, maps.spawn_pool(m4, 1024, rng, min_goal_dist=lo, max_goal_dist=hi), n=500)
        tot.append(g)
        print('  C4 map %dx%d: %.0f gathers/scan (%.2f samples/ray, max %d)' % (m4['width'], m4['height'], g, spr, mx))
    print('C4 pool mean: %.0f gathers/scan -> x1.01 = %.0f per env-step' % (np.mean(tot), 1.01 * np.mean(tot)))
