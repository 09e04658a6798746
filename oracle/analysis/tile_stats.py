"""One-off: fraction of march samples that fall within +-h cells of the origin cell."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, numba
from oracle import oracle as orc
from bench import build_world
from nav_gym_b200 import maps as M

@numba.njit(cache=True)
def stats(dist, W, H, ox, oy, heads, tstop, hs, out, far16):
    n = len(heads)
    tot = 0
    for i in range(n):
        dx = np.float32(np.cos(np.float64(heads[i]))); dy = np.float32(np.sin(np.float64(heads[i])))
        t = np.float32(0)
        while t < tstop:
            px = int(np.float32(dx*t + ox[i])); py = int(np.float32(dy*t + oy[i]))
            if px < 0 or px >= W or py < 0 or py >= H: break
            d = dist[py, px]; tot += 1
            m = max(abs(px - int(ox[i])), abs(py - int(oy[i])))
            for j in range(len(hs)):
                if m < hs[j]: out[j] += 1
            if d >= 16: far16[0] += 1
            if d <= 0: break
            st = np.float32(d*np.float32(0.999))
            if st <= 1: st = np.float32(1)
            t = np.float32(t + st)
    return tot

def run(m, pool, name):
    dist = orc.edt(np.asarray(m['data']) >= 0.1)
    rng = np.random.RandomState(0)
    N = 1000
    rows = pool[rng.randint(len(pool), size=N)]
    ox = np.repeat((rows[:,0]/0.05).astype(np.int32).astype(np.float32), 512)
    oy = np.repeat((rows[:,1]/0.05).astype(np.int32).astype(np.float32), 512)
    heads = (np.tile(orc.beam_table(), N) + np.repeat(rows[:,4], 512)).astype(np.float32)
    hs = np.array([32, 48, 64, 96, 128, 192, 256], np.int64); out = np.zeros(len(hs), np.int64); far = np.zeros(1, np.int64)
    tot = stats(dist, m['width'], m['height'], ox, oy, heads, np.float32(502), hs, out, far)
    print(name, 'steps/ray %.2f' % (tot/len(heads)), ' within +-h:', {int(h): round(o/tot,3) for h,o in zip(hs,out)}, ' d>=16: %.3f' % (far[0]/tot))

m, pool = build_world(0, 8192); run(m, pool, 'indoor cw3')
rng = np.random.RandomState(5)
m2 = M.create_indoor_map(4, 150, rng); run(m2, M.spawn_pool(m2, 4096, rng), 'indoor cw4')
m3 = M.create_outdoor_map(10, 0.7, rng); run(m3, M.spawn_pool(m3, 4096, rng, min_goal_dist=5), 'outdoor400')
m4 = M.create_large_outdoor_map(rng); run(m4, M.spawn_pool(m4, 4096, rng), 'outdoor2000')
