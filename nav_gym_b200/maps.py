"""World generation for the batched simulator: occupancy maps in the style of the reference's
map_generator.py (indoor corridor trees :97-123, outdoor boxes :126-143), the 0.25 m cost-map
of NavGymEnv._sample_map (env.py:309-332) and precomputed spawn pools that stand in for the
reference's per-episode start/goal rejection sampling (env.py:342-383, 748-783) so that
environments can auto-reset on the device.

Reset-time, host-side code (numpy); nothing here is on the per-step path.  Map dicts use the
reference's map_info layout: data int8 [row=y][col=x] with 0 free / 100 occupied, origin,
resolution, width, height.
"""
import ctypes as C

import numpy as np

from . import _lib


# fixture-minting scripts (oracle/make_bench_world.py) route the BFS through the CPU checker so
# that they run without the CUDA library; None = the library's navgym_grid_bfs
_BFS_OVERRIDE = None


def _map_info(occ, resolution=0.05):
    data = np.zeros(occ.shape, np.int8)
    data[occ.astype(bool)] = 100
    return dict(data=data, origin=(0, 0), resolution=resolution, width=occ.shape[1],
                height=occ.shape[0])


def create_outdoor_map(obstacle_number=10, obstacle_width=0.7, rng=None, size=400, border=5):
    """Open square with `obstacle_number` axis-aligned boxes of half-width
    int(10*obstacle_width) cells and a `border`-cell wall (map_generator.py:126-143)."""
    rng = np.random if rng is None else rng
    w = int(10 * obstacle_width)
    occ = np.ones((size, size), np.uint8)
    occ[border:size - border, border:size - border] = 0
    for _ in range(int(obstacle_number)):
        cx = rng.randint(w + 2, size - w - 1)
        cy = rng.randint(w + 2, size - w - 1)
        occ[cx - w:cx + w + 1, cy - w:cy + w + 1] = 1
    return _map_info(np.flipud(occ))


def create_large_outdoor_map(rng=None, size=2000, obstacle_number=250):
    """the outdoor generator at 100 m x 100 m with 250 boxes of half-width
    int(10*U[0.3,1.0]) cells."""
    rng = np.random if rng is None else rng
    occ = np.ones((size, size), np.uint8)
    occ[5:size - 5, 5:size - 5] = 0
    for _ in range(int(obstacle_number)):
        w = int(10 * rng.uniform(0.3, 1.0))
        cx = rng.randint(w + 2, size - w - 1)
        cy = rng.randint(w + 2, size - w - 1)
        occ[cx - w:cx + w + 1, cy - w:cy + w + 1] = 1
    return _map_info(np.flipud(occ))


def create_indoor_map(corridor_width=3, iterations=100, rng=None, cells=100, scale=10):
    """Random corridor tree on a cells x cells grid (map_generator.py:97-123): every new node is
    joined to its L1-nearest tree node (first of equals, :26-35) by an L-shaped corridor of
    half-width `corridor_width` whose elbow is one of the two corners of the nodes' bounding
    box that is not a node, chosen by a coin flip (:57-85); the grid is then upsampled by
    `scale` (100 x 100 -> 1000 x 1000 at 0.05 m) and flipped.  Draws from `rng` in the
    reference's order (x, y, coin per iteration), so `rng=None` under np.random.seed(s), or
    RandomState(s), reproduces the reference's map for that seed bit for bit
    (tests/golden/bench_world.npz, tests/test_host_helpers.py)."""
    rng = np.random if rng is None else rng
    r = int(corridor_width)
    occ = np.ones((cells, cells), np.uint8)
    nodes = [(cells // 2, cells // 2)]
    occ[nodes[0]] = 0
    for _ in range(int(iterations)):
        p = (int(rng.randint(r + 2, cells - r - 1)), int(rng.randint(r + 2, cells - r - 1)))
        arr = np.asarray(nodes)
        q = nodes[int(np.argmin(np.abs(arr[:, 0] - p[0]) + np.abs(arr[:, 1] - p[1])))]
        nodes.append(p)
        occ[p] = 0
        heads = rng.random_sample() >= 0.5
        x1, x2 = sorted((p[0], q[0]))
        y1, y2 = sorted((p[1], q[1]))
        # the nodes sit on the anti-diagonal of their bounding box (x1, y2) / (x2, y1) or on its
        # diagonal; the elbow is a corner of the other diagonal
        anti = (p[0] > q[0] and p[1] < q[1]) or (p[0] < q[0] and p[1] > q[1])
        if anti:
            xc, yc = (x1, y1) if heads else (x2, y2)
        else:
            xc, yc = (x1, y2) if heads else (x2, y1)
        occ[xc - r:xc + r + 1, y1 - r:y2 + r + 1] = 0
        occ[x1 - r:x2 + r + 1, yc - r:yc + r + 1] = 0
    occ = np.kron(occ, np.ones((scale, scale), np.uint8))
    return _map_info(np.flipud(occ))


def cost_map(map_info, new_resolution=0.25, inflate=4):
    """0.25 m planning grid of env.py:312-332: nearest-neighbour downsample, then every cell
    within `inflate` cells (1 m) of an obstacle is blocked (9x9 box filter > 0)."""
    step = int(round(new_resolution / map_info['resolution']))
    occ = (np.asarray(map_info['data'])[::step, ::step] > 0)
    k = inflate
    pad = np.pad(occ, k, mode='reflect')
    csum = np.cumsum(np.cumsum(pad.astype(np.int32), 0), 1)
    csum = np.pad(csum, ((1, 0), (1, 0)))
    n = 2 * k + 1
    H, W = occ.shape
    box = csum[n:n + H, n:n + W] - csum[:H, n:n + W] - csum[n:n + H, :W] + csum[:H, :W]
    data = np.where(box > 0, 100, 0).astype(np.uint8)
    return dict(data=data, origin=map_info['origin'], resolution=new_resolution,
                width=data.shape[1], height=data.shape[0])


def grid_bfs(blocked, start_rc):
    """4-connected geodesic distance (cells) from start over free cells; -1 = unreachable.
    Stands in for pyastar2d.astar_path on the uniform-cost grid of env.py:343-354."""
    if _BFS_OVERRIDE is not None:
        return _BFS_OVERRIDE(blocked, start_rc)
    lib = _lib.load()
    b = np.ascontiguousarray(blocked, np.uint8)
    out = np.empty(b.shape, np.int32)
    lib.navgym_grid_bfs(b.ctypes.data_as(C.c_void_p), b.shape[0], b.shape[1], int(start_rc[0]),
                        int(start_rc[1]), out.ctypes.data_as(C.c_void_p))
    return out


def spawn_pool(map_info, n, rng=None, min_goal_dist=10., max_goal_dist=20., goals_per_start=16,
               max_detour=2.0):
    """(start_x, start_y, goal_x, goal_y, theta) tuples obeying the reference's episode law:
    both ends on free cost-map cells (>= 1 m clearance, env.py:328-332), straight-line
    distance in (min, max) (env.py:379), connected with a path no longer than `max_detour`
    times the straight line (env.py:761, measured here on the 4-connected grid), uniform
    heading (env.py:763).  The discomfort-free-spawn filter (env.py:779-783) needs a lidar
    scan and is applied on the device by BatchedNavGym.filter_spawn_pool."""
    rng = np.random if rng is None else rng
    cm = cost_map(map_info)
    blocked = cm['data'] > 0
    rows, cols = np.where(~blocked)
    if len(rows) == 0:
        return np.zeros((0, 5))
    res, (ox, oy) = cm['resolution'], cm['origin']
    out = []
    tries = 0
    while sum(len(o) for o in out) < n and tries < 20 * (n // goals_per_start + 1):
        tries += 1
        s = rng.randint(len(rows))
        sr, sc = rows[s], cols[s]
        geo = grid_bfs(blocked, (sr, sc)) * res
        sx, sy = (sc + 0.5) * res + ox, (sr + 0.5) * res + oy
        gx, gy = (cols + 0.5) * res + ox, (rows + 0.5) * res + oy
        d = np.hypot(gx - sx, gy - sy)
        g = geo[rows, cols]
        ok = np.where((d > min_goal_dist) & (d < max_goal_dist) & (g >= 0) & (g <= max_detour * d))[0]
        if len(ok) == 0:
            continue
        pick = ok[rng.randint(len(ok), size=min(goals_per_start, len(ok)))]
        th = rng.uniform(0, 2 * np.pi, len(pick))
        out.append(np.column_stack([np.full(len(pick), sx), np.full(len(pick), sy), gx[pick], gy[pick], th]))
    if not out:
        return np.zeros((0, 5))
    pool = np.concatenate(out)[:n]
    return pool[rng.permutation(len(pool))]


def spawn_pedestrians(map_info, robot_xy, num_peds, rng=None, v_pref_range=(0.0, 0.6),
                      has_legs_ratio=0.5, min_robot_dist=4.0, min_goal_dist=10.0, trunk_radius=0.3):
    """Pedestrian state rows (layout include/navgym_b200.h, NAVGYM_PED_F floats each) for one
    environment, following the reference's spawn law (env.py:786-806): start on a free
    cost-map cell at least 4 m from the robot, a goal at least 10 m away that is connected to
    it, preferred speed U[0, 0.6], legs with probability 0.5.  The two waypoints are the start
    and the goal (the reference walks A*-waypoints towards the goal and then re-plans,
    env.py:633-680; the scripted stand-in walks back and forth)."""
    rng = np.random if rng is None else rng
    cm = cost_map(map_info)
    blocked = cm['data'] > 0
    rows, cols = np.where(~blocked)
    res, (ox, oy) = cm['resolution'], cm['origin']
    out = np.zeros((num_peds, _lib.PED_F), np.float32)
    if len(rows) == 0:
        return out
    xs, ys = (cols + 0.5) * res + ox, (rows + 0.5) * res + oy
    for p in range(num_peds):
        for _ in range(100):
            i = rng.randint(len(rows))
            if np.hypot(xs[i] - robot_xy[0], ys[i] - robot_xy[1]) < min_robot_dist:
                continue
            geo = grid_bfs(blocked, (rows[i], cols[i]))
            d = np.hypot(xs - xs[i], ys - ys[i])
            ok = np.where((d > min_goal_dist) & (geo[rows, cols] >= 0))[0]
            if len(ok) == 0:
                ok = np.where((d > 0.3 * min_goal_dist) & (geo[rows, cols] >= 0))[0]
            if len(ok):
                j = ok[rng.randint(len(ok))]
                break
        else:
            j = i
        th = rng.uniform(0, 2 * np.pi)
        out[p, :4] = (xs[i], ys[i], th, rng.uniform(*v_pref_range))
        out[p, 4:8] = (xs[i], ys[i], xs[j], ys[j])
        out[p, 8] = 1.0
        out[p, 12] = float(rng.random_sample() < has_legs_ratio)
        out[p, 13] = trunk_radius
    return out


def goal_fields(map_info, num_goals=32, rng=None, new_resolution=0.25, inflate=4):
    """Geodesic distance fields for the pedestrians' route following (navgym_peds_plan): pick
    `num_goals` free cells of the cost map (env.py:312-332, where the reference samples
    pedestrian goals, :375-377) and BFS from each over the cost map's free cells (the
    uniform-cost grid the reference's A* runs on, :343-351).  Cells inside the 1 m inflation
    band are then filled outward from the reachable region, so that a pedestrian that strayed
    next to a wall still has a downhill direction.  Returns (fields uint16 [G, H, W] with 65535 =
    unreachable, goals float64 [G, 2] world xy of the goal cells' centres, cost map dict)."""
    rng = np.random if rng is None else rng
    cm = cost_map(map_info, new_resolution, inflate)
    blocked = cm['data'] > 0
    step = int(round(new_resolution / map_info['resolution']))
    raw_free = ~(np.asarray(map_info['data'])[::step, ::step] > 0)
    rows, cols = np.where(~blocked)
    if len(rows) == 0:
        raise ValueError('cost map has no free cell')
    # goals inside the largest connected region, so that most pairs are mutually reachable
    seed_geo = grid_bfs(blocked, (rows[len(rows) // 2], cols[len(rows) // 2]))
    ok = np.where(seed_geo[rows, cols] >= 0)[0]
    if len(ok) < max(8, len(rows) // 4):
        best = ok
        for _ in range(8):
            i = rng.randint(len(rows))
            cand = np.where(grid_bfs(blocked, (rows[i], cols[i]))[rows, cols] >= 0)[0]
            if len(cand) > len(best):
                best = cand
        ok = best
    pick = ok[rng.choice(len(ok), size=num_goals, replace=len(ok) < num_goals)]
    H, W = blocked.shape
    fields = np.empty((num_goals, H, W), np.uint16)
    for g, i in enumerate(pick):
        d = grid_bfs(blocked, (rows[i], cols[i])).astype(np.int64)
        for _ in range(inflate + 2):
            p = np.pad(d, 1, constant_values=-1)
            nb = np.stack([p[1:-1, :-2], p[1:-1, 2:], p[:-2, 1:-1], p[2:, 1:-1]])
            best = np.where(nb >= 0, nb, np.iinfo(np.int64).max).min(0)
            fill = (d < 0) & raw_free & (best < np.iinfo(np.int64).max)
            d = np.where(fill, best + 1, d)
        fields[g] = np.where(d < 0, 65535, np.minimum(d, 65534)).astype(np.uint16)
    res, (ox, oy) = cm['resolution'], cm['origin']
    goals = np.column_stack([(cols[pick] + 0.5) * res + ox, (rows[pick] + 0.5) * res + oy]).astype(np.float64)
    return fields, goals, cm
