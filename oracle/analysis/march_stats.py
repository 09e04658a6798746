"""One-off analysis (uses the CPU oracle, so it lives under oracle/): distribution of march steps and unit-step runs on the bench world."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, numba
from oracle import oracle as orc
from bench import build_world

@numba.njit(cache=True)
def stats(dist, W, H, ox, oy, heads, tstop):
    n = len(heads)
    steps = np.zeros(n, np.int32); unit = np.zeros(n, np.int32); maxrun = np.zeros(n, np.int32)
    for i in range(n):
        dx = np.float32(np.cos(np.float64(heads[i]))); dy = np.float32(np.sin(np.float64(heads[i])))
        t = np.float32(0); run = 0
        while t < tstop:
            px = int(np.float32(dx*t + ox[i])); py = int(np.float32(dy*t + oy[i]))
            if px < 0 or px >= W or py < 0 or py >= H: break
            d = dist[py, px]; steps[i] += 1
            if d <= 0: break
            st = np.float32(d*np.float32(0.999))
            if st <= 1: 
                st = np.float32(1); unit[i] += 1; run += 1
                if run > maxrun[i]: maxrun[i] = run
            else: run = 0
            t = np.float32(t + st)
    return steps, unit, maxrun

m, pool = build_world(0, 8192)
dist = orc.edt(np.asarray(m['data']) >= 0.1)
rng = np.random.RandomState(0)
N = 2000
rows = pool[rng.randint(len(pool), size=N)]
ox = np.repeat((rows[:,0]/0.05).astype(np.int32).astype(np.float32), 512)
oy = np.repeat((rows[:,1]/0.05).astype(np.int32).astype(np.float32), 512)
heads = (np.tile(orc.beam_table(), N) + np.repeat(rows[:,4], 512)).astype(np.float32)
steps, unit, maxrun = stats(dist, 1000, 1000, ox, oy, heads, np.float32(502))
print('rays', len(steps), 'mean steps', steps.mean(), 'p50', np.percentile(steps,50), 'p90', np.percentile(steps,90), 'p99', np.percentile(steps,99), 'max', steps.max())
print('unit-step fraction of steps', unit.sum()/steps.sum())
for thr in (2,3,4,6,8,12,16):
    sel = maxrun >= thr
    print('maxrun>=%2d: %.3f%% of rays, holding %.1f%% of all steps' % (thr, 100*sel.mean(), 100*steps[sel].sum()/steps.sum()))
s = steps.reshape(N, 512)
print('per-env: mean of max ray steps', s.max(1).mean(), ' per-env sum', s.sum(1).mean())
# lane sums for one-warp-per-env with 16 beams/lane (k = lane + 32 i)
ls = s.reshape(N, 16, 32).sum(1)
print('warp/env: mean lane-sum', ls.mean(), 'mean max-lane-sum', ls.max(1).mean())
# 4 warps per env, 4 beams per lane: k = lane + 32*(warp + 4 i)
l4 = s.reshape(N, 4, 4, 32).sum(1)   # [N, warp, lane]
print('4 warps/env: mean lane-sum', l4.mean(), 'mean max over lanes per warp', l4.max(2).mean(), 'mean max per env', l4.max(2).max(1).mean())
# cap steps at c (rest deferred)
for c in (16, 24, 32, 48):
    sc = np.minimum(s, c); l4c = sc.reshape(N,4,4,32).sum(1)
    print('cap %d: deferred rays %.2f%%, mean max lane-sum per warp %.1f (mean lane-sum %.1f)' % (c, 100*(s>c).mean(), l4c.max(2).mean(), l4c.mean()))

# ---- list-scheduling simulation: beams handed out in order to the first free lane
import heapq
def makespan(job, lanes):
    h = [0]*lanes
    heapq.heapify(h)
    for j in job:
        t = heapq.heappop(h); heapq.heappush(h, t + j)
    return max(h)
s1 = np.maximum(s - 1, 1)   # first sample shared per env
for lanes in (32, 64, 128, 256):
    dyn = np.mean([makespan(s1[i], lanes) for i in range(300)])
    # longest-first variant (oracle knowledge) for reference
    lpt = np.mean([makespan(np.sort(s1[i])[::-1], lanes) for i in range(300)])
    wpe = lanes // 32
    stat = s1[:300].reshape(300, 16 // wpe, wpe, 32).sum(1).max(2).max(1).mean()
    print('lanes %3d: ideal %.1f  dynamic in-order %.1f  longest-first %.1f  static (current) %.1f' % (
        lanes, s1[:300].sum(1).mean() / lanes, dyn, lpt, stat))

# ---- lockstep groups of 32 adjacent beams (warp-uniform params, no per-lane refill)
g = s1.reshape(N, 16, 32)
print('lockstep 32 adjacent beams: mean steps/ray %.2f, mean of group max %.2f -> SIMT efficiency %.2f' % (
    g.mean(), g.max(2).mean(), g.mean() / g.max(2).mean()))
for cap in (8, 12, 16, 24):
    # two-phase: lockstep up to `cap` iterations, leftovers compacted and dealt dynamically
    main = np.minimum(g, cap); rest = np.maximum(g - cap, 0)
    it_main = np.minimum(g.max(2), cap).mean()
    frac_left = (rest > 0).mean(); steps_left = rest.sum() / g.sum()
    print(' cap %2d: main iterations/group %.2f (eff %.2f), rays continuing %.1f%%, steps left %.1f%%' % (
        cap, it_main, main.mean() / it_main, 100 * frac_left, 100 * steps_left))
