"""One-off: tail-phase iterations per warp of the step kernel's march under different orders of
dealing the surviving beams to lanes (in beam order = the kernel today; longest remaining first =
the ideal; buckets by the size of the beam's last head-phase step = a proxy the kernel has)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, numba
from oracle import oracle as orc
from bench import build_world

@numba.njit(cache=True)
def beam_profile(dist, W, H, ox, oy, heads, tstop, HS, remaining, last_step):
    d0 = dist[int(oy), int(ox)]
    for k in range(512):
        remaining[k] = 0; last_step[k] = 0.0
    if d0 <= 0: return
    t1 = np.float32(max(np.float32(d0 * np.float32(0.999)), np.float32(1.0)))
    for k in range(512):
        dx = np.float32(np.cos(np.float64(heads[k]))); dy = np.float32(np.sin(np.float64(heads[k])))
        t = t1; n = 0; alive = True; ls = np.float32(0)
        while alive:
            cx = int(np.float32(dx * t + ox)); cy = int(np.float32(dy * t + oy))
            if cx < 0 or cx >= W or cy < 0 or cy >= H: break
            d = dist[cy, cx]; n += 1
            if d <= 0: break
            st = max(np.float32(d * np.float32(0.999)), np.float32(1.0))
            if n == HS: ls = st
            t = np.float32(t + st)
            if not (t < tstop): break
        remaining[k] = max(n - HS, 0)
        last_step[k] = ls

@numba.njit(cache=True)
def tail_iters(order, remaining):
    # two warps; warp w takes entries w, w+2, ... of `order`; lanes refill as beams end
    tot = 0
    for w in range(2):
        mine = order[w::2]
        n = len(mine)
        lane = np.zeros(32, np.int64)
        nxt = 0
        for l in range(32):
            if nxt < n: lane[l] = remaining[mine[nxt]]; nxt += 1
        it = 0
        while True:
            busy = False
            for l in range(32):
                if lane[l] > 0:
                    busy = True; lane[l] -= 1
                    if lane[l] == 0 and nxt < n:
                        lane[l] = remaining[mine[nxt]]; nxt += 1
            if not busy: break
            it += 1
        tot = max(tot, it)  # the CTA waits for the slower warp
    return tot

m, pool = build_world(0, 8192)
dist = orc.edt(np.asarray(m['data']) >= 0.1)
rng = np.random.RandomState(0)
N = 400
rows = pool[rng.randint(len(pool), size=N)]
bt = orc.beam_table()
res = {}
rem = np.zeros(512, np.int64); ls = np.zeros(512, np.float32)
rem0 = np.zeros(512, np.int64); ls0 = np.zeros(512, np.float32)
INC = 0.0122718437
for r in rows:
    ox = np.float32(int(r[0] / 0.05)); oy = np.float32(int(r[1] / 0.05))
    heads = (bt + np.float32(r[4])).astype(np.float32)
    beam_profile(dist, m['width'], m['height'], ox, oy, heads, np.float32(502), 4, rem, ls)
    surv = np.where(rem > 0)[0]
    # the previous step's scan: the robot came from 0.1 m behind, turned by up to 0.128 rad
    v, w = rng.uniform(0, 0.5), rng.uniform(-0.64, 0.64)
    th0 = r[4] - w * 0.2
    ox0 = np.float32(int((r[0] - v * 0.2 * np.cos(r[4])) / 0.05)); oy0 = np.float32(int((r[1] - v * 0.2 * np.sin(r[4])) / 0.05))
    beam_profile(dist, m['width'], m['height'], ox0, oy0, (bt + np.float32(th0)).astype(np.float32), np.float32(502), 4, rem0, ls0)
    shift = int(np.rint((r[4] - th0) / INC))
    pred = rem0[(np.arange(512) + shift) % 512]          # beam k now looked along old beam k + shift
    orders = {
        'history: longest predicted first': surv[np.argsort(-pred[surv], kind='stable')],
        'history: predicted >= 16 first': np.concatenate([surv[pred[surv] >= 16], surv[pred[surv] < 16]]),
        'history: predicted >= 8 first': np.concatenate([surv[pred[surv] >= 8], surv[pred[surv] < 8]]),
        'history, no shift: >= 16 first': np.concatenate([surv[rem0[surv] >= 16], surv[rem0[surv] < 16]]),
        'beam order (today)': surv,
        'longest first (ideal)': surv[np.argsort(-rem[surv], kind='stable')],
        'last step < 4 cells first': np.concatenate([surv[ls[surv] < 4], surv[ls[surv] >= 4]]),
        'last step < 8 cells first': np.concatenate([surv[ls[surv] < 8], surv[ls[surv] >= 8]]),
        'last step < 16 cells first': np.concatenate([surv[ls[surv] < 16], surv[ls[surv] >= 16]]),
        'ascending last step': surv[np.argsort(ls[surv], kind='stable')],
    }
    for k, o in orders.items():
        res.setdefault(k, []).append(tail_iters(o.astype(np.int64), rem))
    res.setdefault('ideal = samples / 64', []).append(int(np.ceil(rem.sum() / 64)))
    res.setdefault('longest beam', []).append(int(rem.max()))
for k, v in res.items():
    print('%-28s mean %.1f  p90 %.0f  max %d' % (k, np.mean(v), np.percentile(v, 90), np.max(v)))
