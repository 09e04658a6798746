"""Per-call latency of the level-1 binding: the reference's three native
calls per scan -- calc_range_many on 512 rays, render_contours_in_lidar on 15 footprints,
render_agents_in_lidar on 15 legged agents -- through nav_gym_b200.natives (numpy in / numpy out)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from nav_gym_b200 import maps, natives

rng = np.random.RandomState(0)
m = maps.create_indoor_map(3, 100, rng)
occ = np.ascontiguousarray(np.asarray(m['data']) >= 0.1)
rm = natives.PyRayMarching(natives.PyOMap(occ), float(occ.size))
ins = np.zeros((512, 3), np.float32)
free = np.argwhere(~occ)
cy, cx = free[len(free) // 2]
ins[:, 0], ins[:, 1] = cx, cy
ins[:, 2] = np.linspace(-3.141592, 3.141592 - 0.0122718463, 512)
outs = np.zeros(512, np.float32)
angles = ins[:, 2].copy()
xy = np.array([cx * 0.05, cy * 0.05], np.float32)
polys = [[[xy[0] + 2 + 0.3 * i, xy[1] + 1], [xy[0] + 2.4 + 0.3 * i, xy[1] + 1], [xy[0] + 2.4 + 0.3 * i, xy[1] + 1.4],
          [xy[0] + 2 + 0.3 * i, xy[1] + 1.4]] for i in range(15)]
flat = natives.flatten_contours(polys)
cm = natives.CMap2D()
cm.set_resolution(1.)
agents = [natives.CSimAgent(np.array([xy[0] - 2 - 0.2 * i, xy[1] + 0.5, 0.3], np.float32),
                            np.array([0.4, 0.1, 0.2], np.float32), np.zeros(2, np.float32)) for i in range(15)]


def timeit(fn, n=300):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


ranges = np.full(512, 25.0, np.float32)
out = {"calc_range_many_512_us": timeit(lambda: rm.calc_range_many(ins, outs)),
       "render_contours_15_us": timeit(lambda: natives.render_contours_in_lidar(ranges, angles, flat, xy)),
       "render_agents_15_us": timeit(lambda: cm.render_agents_in_lidar(ranges, angles, agents, xy))}
out["scan_of_one_agent_us"] = sum(out.values())
print(json.dumps(out))
