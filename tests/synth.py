"""Seeded synthetic worlds for parity tests (maps in the style of the reference's
map_generator.py:97-143, restated — the reference itself is not on the GPU box)."""
import numpy as np


def outdoor_map(rng, size=400, n_obs=10, border=5):
    m = np.ones((size, size), np.int8)
    m[border:size - border, border:size - border] = 0
    for _ in range(n_obs):
        w = int(10 * rng.uniform(0.3, 1.0))
        cx, cy = rng.randint(w + 2, size - w - 1, 2)
        m[cx - w:cx + w + 1, cy - w:cy + w + 1] = 1
    data = np.zeros((size, size), np.int8)
    data[m == 1] = 100
    return dict(data=data, origin=(0, 0), resolution=0.05, width=size, height=size)


def indoor_map(rng, cells=100, scale=10, corridor=3, iterations=100):
    m = np.ones((cells, cells), np.int8)
    tree = [(cells // 2, cells // 2)]
    m[tree[0]] = 0
    for _ in range(iterations):
        p = (rng.randint(corridor + 2, cells - corridor - 1), rng.randint(corridor + 2, cells - corridor - 1))
        q = min(tree, key=lambda n: abs(n[0] - p[0]) + abs(n[1] - p[1]))
        tree.append(p)
        x1, x2 = sorted((p[0], q[0]))
        y1, y2 = sorted((p[1], q[1]))
        xc = p[0] if rng.rand() < 0.5 else q[0]
        yc = q[1] if xc == p[0] else p[1]
        m[xc - corridor:xc + corridor + 1, y1 - corridor:y2 + corridor + 1] = 0
        m[x1 - corridor:x2 + corridor + 1, yc - corridor:yc + corridor + 1] = 0
    m = np.kron(m, np.ones((scale, scale), np.int8))
    data = np.zeros(m.shape, np.int8)
    data[m == 1] = 100
    data = np.flipud(data).copy()
    return dict(data=data, origin=(0, 0), resolution=0.05, width=m.shape[1], height=m.shape[0])


def free_poses(rng, m, n, clearance_cells, edt):
    """n random (x, y) in free space at least clearance_cells from any obstacle."""
    ys, xs = np.where(edt >= clearance_cells)
    idx = rng.randint(0, len(xs), n)
    res = m['resolution']
    x = (xs[idx] + rng.uniform(0, 1, n)) * res + m['origin'][0]
    y = (ys[idx] + rng.uniform(0, 1, n)) * res + m['origin'][1]
    return np.column_stack([x, y])


def random_geometry(rng, B, start, max_disc, max_seg, spread=6.0):
    """discs / box-footprint segments scattered around each env's robot."""
    discs = np.zeros((B, max_disc, 3), np.float32)
    segs = np.zeros((B, max_seg, 4), np.float32)
    ndisc = rng.randint(0, max_disc + 1, B).astype(np.int32)
    nbox = rng.randint(0, max_seg // 4 + 1, B)
    nseg = (4 * nbox).astype(np.int32)
    for e in range(B):
        c = start[e] + rng.uniform(-spread, spread, (max_disc, 2))
        discs[e, :, :2] = c
        discs[e, :, 2] = rng.choice([0.03, 0.3], max_disc)
        for b in range(nbox[e]):
            cx, cy = start[e] + rng.uniform(-spread, spread, 2)
            th = rng.uniform(0, 2 * np.pi)
            fp = np.array([[0.22, 0.19], [-0.22, 0.19], [-0.22, -0.19], [0.22, -0.19]])
            co, si = np.cos(th), np.sin(th)
            w = np.column_stack([co * fp[:, 0] - si * fp[:, 1] + cx, si * fp[:, 0] + co * fp[:, 1] + cy])
            segs[e, 4 * b:4 * b + 4] = np.concatenate([w, np.roll(w, -1, axis=0)], axis=1)
    return discs, ndisc, segs, nseg


def her_batch(seed=5, n=6000):
    """Stored observations for the HER entry points (compute_rewards / compute_terminals,
    env.py:491-589): float32 rows [scan(512) | prev_pose pose vel yaw] and goals, a mix of rows
    clear of every threshold, rows in discomfort (some beams between the crash and the discomfort
    threshold of their angle), crashed rows and rows at the goal.  The same seed gives the same
    batch in oracle/make_golden_her.py (which pushes it through the reference's own functions)
    and in the tests; thresholds are passed in because they come from the implementation under
    test / the reference."""
    rng = np.random.RandomState(seed)
    scan = rng.uniform(2.0, 25.0, (n, 512)).astype(np.float32)
    kind = rng.randint(0, 4, n)                      # 0 clear, 1 discomfort, 2 crash, 3 clear + at goal
    frac = rng.uniform(0.02, 0.98, (n, 512)).astype(np.float32)
    pick = rng.uniform(0, 1, (n, 512)) < 0.02        # ~10 beams per row carry the close return
    pose = rng.uniform(5, 45, (n, 2))
    prev = pose + rng.uniform(-0.1, 0.1, (n, 2))
    vel = np.column_stack([rng.uniform(0, 0.5, n), rng.uniform(-0.64, 0.64, n)])
    yaw = rng.uniform(-np.pi, np.pi, n)
    goal = pose + rng.uniform(-15, 15, (n, 2))
    at_goal = kind == 3
    goal[at_goal] = pose[at_goal] + rng.uniform(-0.3, 0.3, (int(at_goal.sum()), 2))
    tail = np.column_stack([prev, pose, vel, yaw]).astype(np.float32)
    return dict(scan=scan, kind=kind, frac=frac, pick=pick, tail=tail, goal=goal.astype(np.float32))


def her_rows(batch, thr, dthr):
    """The observation rows of her_batch() for the given crash / discomfort threshold vectors."""
    scan = batch['scan'].copy()
    thr, dthr = np.asarray(thr, np.float32), np.asarray(dthr, np.float32)
    between = (thr + batch['frac'] * (dthr - thr)).astype(np.float32)     # inside the discomfort band
    below = (thr * batch['frac']).astype(np.float32)                       # inside the footprint
    d = (batch['kind'] == 1)[:, None] & batch['pick']
    c = (batch['kind'] == 2)[:, None] & batch['pick']
    scan[d] = between[d]
    scan[c] = below[c]
    return np.concatenate([scan, batch['tail']], axis=1).astype(np.float32)
