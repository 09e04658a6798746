"""One-off analysis (uses the CPU checker, so it lives under oracle/): how many L2 round trips the
march of a beam needs when N lanes fetch the cells at t, t + delta, ... t + (N-1) delta together and
the true march is then walked through the fetched cells (step_kernel.cuh, tail regime B), against one
round trip per sample; lookup among 1 or 3 candidate lanes.  Bench world, 1500 poses x 512 beams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, numba
from oracle import oracle as orc
from nav_gym_b200 import worlds

@numba.njit(cache=True)
def sim(dist, W, H, ox, oy, heads, tstop, N, delta, head_steps, ncand):
    n = len(heads)
    steps = np.zeros(n, np.int32); rounds = np.zeros(n, np.int32)
    cells = np.zeros(64, np.int64)
    for i in range(n):
        dx = np.float32(np.cos(np.float64(heads[i]))); dy = np.float32(np.sin(np.float64(heads[i])))
        t = np.float32(0)
        # head: plain sequential steps
        done = False
        k = 0
        while True:
            # round start at t: fetch N cells
            if k >= head_steps:
                rounds[i] += 1
                for j in range(N):
                    tj = np.float32(t + np.float32(j * delta))
                    px = int(np.float32(dx*tj + ox[i])); py = int(np.float32(dy*tj + oy[i]))
                    if px < 0 or px >= W or py < 0 or py >= H: cells[j] = -1
                    else: cells[j] = py * W + px
                t0 = t
            first = True
            while True:
                if not (t < tstop): done = True; break
                px = int(np.float32(dx*t + ox[i])); py = int(np.float32(dy*t + oy[i]))
                if px < 0 or px >= W or py < 0 or py >= H: done = True; break
                c = py * W + px
                if k >= head_steps and not first:
                    # lookup among fetched
                    jj = int(round((t - t0) / delta))
                    found = False
                    for qi in range(ncand):
                        q = jj if qi == 0 else (jj - 1 if qi == 1 else jj + 1)
                        if q >= 0 and q < N and cells[q] == c: found = True
                    if not found: break   # new round from this t
                d = dist[py, px]; steps[i] += 1; k += 1
                first = False
                if d <= 0: done = True; break
                st = np.float32(d*np.float32(0.999))
                if st < 1: st = np.float32(1)
                t = np.float32(t + st)
                if k < head_steps + 1 and k >= head_steps: break
                if k < head_steps: 
                    rounds[i] += 1
                    first = True
            if done: break
    return steps, rounds

m, pool = worlds.load_bench_world()
dist = orc.edt(np.asarray(m['data']) >= 0.1)
rng = np.random.RandomState(0)
NE = 1500
rows = pool[rng.randint(len(pool), size=NE)]
ox = np.repeat((rows[:,0]/0.05).astype(np.int32).astype(np.float32), 512)
oy = np.repeat((rows[:,1]/0.05).astype(np.int32).astype(np.float32), 512)
heads = (np.tile(orc.beam_table(), NE) + np.repeat(rows[:,4], 512)).astype(np.float32)
for (N, delta, nc) in ((8, 1.0, 1), (8, 1.0, 3), (16, 1.0, 1), (16,1.0,3), (32, 1.0, 1), (32, 1.0, 3)):
    steps, rounds = sim(dist, 1000, 1000, ox, oy, heads, np.float32(502), N, delta, 4, nc)
    s = steps.reshape(NE, 512); r = rounds.reshape(NE, 512)
    long = s > 24
    print('cand=%d N=%2d delta=%.2f: mean samples %.2f, mean roundtrips %.2f | long beams (>24 samples, %.2f%%): samples %.1f roundtrips %.1f | per-env max samples %.1f -> max roundtrips %.1f ; global max %d -> %d' % (
        nc, N, delta, s.mean(), r.mean(), 100*long.mean(), s[long].mean(), r[long].mean(), s.max(1).mean(), r.max(1).mean(), s.max(), r.max()))
