"""One-off: distinct 128-byte lines per warp-level EDT gather of the step kernel's march
(head phase = 4 lockstep samples of 32 adjacent beams, tail = ballot-dealt survivors), for the
row-major float32 EDT and for tiled layouts.  Emulates the kernel's lane assignment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, numba
from oracle import oracle as orc
from bench import build_world

@numba.njit(cache=True)
def line_id(cx, cy, mode, W):
    if mode == 0:   # row-major: 32 cells in x
        return cy * 4096 + (cx >> 5)
    if mode == 1:   # 8 x 4 tiles
        return (cy >> 2) * 4096 + (cx >> 3)
    if mode == 2:   # 4 x 8 tiles
        return (cy >> 3) * 4096 + (cx >> 2)
    return (cy >> 1) * 4096 + (cx >> 4)  # 16 x 2

@numba.njit(cache=True)
def count_lines(cxs, cys, act, mode, W):
    ids = np.empty(32, np.int64); n = 0
    for l in range(32):
        if act[l]:
            v = line_id(cxs[l], cys[l], mode, W)
            dup = False
            for q in range(n):
                if ids[q] == v: dup = True; break
            if not dup:
                ids[n] = v; n += 1
    return n

@numba.njit(cache=True)
def sim(dist, W, H, ox, oy, heads, tstop, nmodes, out_head, out_tail, req, HS):
    # one env: ox, oy scalars; heads[512]
    dx = np.empty(512, np.float32); dy = np.empty(512, np.float32)
    for k in range(512):
        dx[k] = np.float32(np.cos(np.float64(heads[k]))); dy[k] = np.float32(np.sin(np.float64(heads[k])))
    d0 = dist[int(oy), int(ox)]
    if d0 <= 0: return
    t1 = np.float32(max(np.float32(d0 * np.float32(0.999)), np.float32(1.0)))
    t = np.full(512, t1, np.float32)
    alive = np.ones(512, np.bool_)
    cxs = np.zeros(32, np.int64); cys = np.zeros(32, np.int64); act = np.zeros(32, np.bool_)
    # head: warp w (0/1), beam index i (0..7): beams 64*i + 32*w + lane
    for i in range(8):
        for w in range(2):
            for st in range(HS):
                anyact = False
                for l in range(32):
                    k = 64 * i + 32 * w + l
                    act[l] = False
                    if alive[k]:
                        cx = int(np.float32(dx[k] * t[k] + ox)); cy = int(np.float32(dy[k] * t[k] + oy))
                        if cx < 0 or cx >= W or cy < 0 or cy >= H:
                            alive[k] = False; continue
                        cxs[l] = cx; cys[l] = cy; act[l] = True; anyact = True
                        d = dist[cy, cx]
                        if d <= 0: alive[k] = False; continue
                        tn = np.float32(t[k] + max(np.float32(d * np.float32(0.999)), np.float32(1.0)))
                        if not (tn < tstop): alive[k] = False
                        t[k] = tn
                if anyact:
                    req[0] += 1
                    for m in range(nmodes): out_head[m] += count_lines(cxs, cys, act, m, W)
    # survivors list in the kernel's order: rounds r=0,1 (4 beams each), j, then warp-interleaved...
    lst = np.empty(512, np.int64); n = 0
    for r in range(2):
        for j in range(4):
            for w in range(2):       # atomicAdd order between warps is arbitrary; take warp 0 first
                for l in range(32):
                    k = 64 * (4 * r + j) + 32 * w + l
                    if alive[k]:
                        lst[n] = k; n += 1
    for w in range(2):
        cur = np.full(32, -1, np.int64)
        nxt = 32
        for l in range(32):
            idx = w + 2 * l
            if idx < n: cur[l] = lst[idx]
        while True:
            anyact = False
            for l in range(32):
                act[l] = False
                k = cur[l]
                if k >= 0: anyact = True
            if not anyact: break
            fins = 0
            realreq = False
            for l in range(32):
                k = cur[l]
                if k < 0: continue
                cx = int(np.float32(dx[k] * t[k] + ox)); cy = int(np.float32(dy[k] * t[k] + oy))
                fin = False
                if cx < 0 or cx >= W or cy < 0 or cy >= H:
                    fin = True
                else:
                    cxs[l] = cx; cys[l] = cy; act[l] = True; realreq = True
                    d = dist[cy, cx]
                    if d <= 0: fin = True
                    else:
                        tn = np.float32(t[k] + max(np.float32(d * np.float32(0.999)), np.float32(1.0)))
                        t[k] = tn
                        if not (tn < tstop): fin = True
                if fin:
                    idx = w + 2 * (nxt + fins); fins += 1
                    cur[l] = lst[idx] if idx < n else -1
            nxt += fins
            if realreq:
                req[1] += 1
                for m in range(nmodes): out_tail[m] += count_lines(cxs, cys, act, m, W)

m, pool = build_world(0, 8192)
dist = orc.edt(np.asarray(m['data']) >= 0.1)
rng = np.random.RandomState(0)
N = 400
rows = pool[rng.randint(len(pool), size=N)]
nm = 4
bt = orc.beam_table()
names = ['row-major 32x1', 'tiled 8x4', 'tiled 4x8', 'tiled 16x2']
for HS in (2, 4, 6, 8, 12, 16, 24):
    oh = np.zeros(nm, np.int64); ot = np.zeros(nm, np.int64); req = np.zeros(2, np.int64)
    for r in rows:
        ox = np.float32(int(r[0] / 0.05)); oy = np.float32(int(r[1] / 0.05))
        heads = (bt + np.float32(r[4])).astype(np.float32)
        sim(dist, m['width'], m['height'], ox, oy, heads, np.float32(502), nm, oh, ot, req, HS)
    print('HEAD_STEPS %d requests/env: head %.1f tail %.1f' % (HS, req[0] / N, req[1] / N))
    for i in range(2):
        print('   %-16s lines/request: head %.2f  tail %.2f   lines/env %.0f' % (names[i], oh[i] / req[0], ot[i] / req[1], (oh[i] + ot[i]) / N))
