"""Runs the reference's own nav_gym_env/env.py UNMODIFIED, inside the build container only.

TEST INFRASTRUCTURE ONLY.  /root/reference is not present on the GPU box, so nothing that
runs there imports this module; its one customer is oracle/make_golden.py, which mints the
fixtures under tests/golden/.

env.py imports five packages that are not installable here: gym, range_libc,
CMap2D (pymap2d), pose2d, pyastar2d; and loads human_policy.pth, a blob missing from the
checkout.  This module puts functional stand-ins into sys.modules BEFORE importing env.py:

  gym          -> nav_gym_b200.gym_shim (the product's registry/spaces shim)
  range_libc   -> PyOMap / PyRayMarching backed by the C oracle (canonical App. B.1)
  CMap2D       -> flatten_contours / render_contours_in_lidar / CMap2D / CSimAgent backed by
                  the C oracle (canonical App. B.2 / B.3)
  pose2d       -> inverse_pose2d / apply_tf_to_vel (App. B.4)
  pyastar2d    -> astar_path, 4-connected A* (reset path only)
  torch.load   -> returns seeded random-init HumanPolicy weights (the .pth is missing)

Every first-party line of env.py / keti_robot.py / human.py / utils.py / map_generator.py
executes as written.  A recorder captures, per _compute_scan call, what crossed the native
boundary so that a trace can be replayed against the oracle and the CUDA path.
"""
import heapq
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from oracle import oracle as orc  # noqa: E402

REF_SRC = '/root/reference/nav_gym/src'


class Recorder(object):
    def __init__(self):
        self.scans = []   # finalized scan records
        self.cur = {}
        # optional transcript of every call that crosses the native boundary, in order, with the
        # arguments as the reference passed them and the result the stand-in produced
        # (oracle/make_golden_native_calls.py); None = off
        self.calls = None

    def log(self, **rec):
        if self.calls is not None:
            self.calls.append(rec)

    def finalize(self, ranges_after, discs):
        rec = self.cur
        rec['discs'] = discs
        rec['ranges_after'] = ranges_after.copy()
        rec.setdefault('segs', np.zeros((0, 4), np.float32))
        rec['noise'] = None
        self.scans.append(rec)
        self.cur = {}

    def clear(self):
        self.scans = []
        self.cur = {}


REC = Recorder()


# ---------------------------------------------------------------- range_libc ------------
class PyOMap(object):
    def __init__(self, arr):
        assert arr.dtype == np.bool_ and arr.flags.c_contiguous
        self.arr = arr


class PyRayMarching(object):
    def __init__(self, omap, max_range):
        self.max_range = float(max_range)
        self.dist = orc.edt(omap.arr)
        REC.log(fn='PyRayMarching', occ=omap.arr.copy(), max_range=float(max_range))

    def calc_range_many(self, ins, outs):
        assert ins.dtype == np.float32 and outs.dtype == np.float32
        assert ins.flags.c_contiguous and outs.flags.c_contiguous
        r, hits = orc.calc_range_many(self.dist, ins, self.max_range, want_hits=True)
        outs[:] = r
        REC.log(fn='calc_range_many', ins=ins.copy(), outs=outs.copy())
        REC.cur['ins'] = ins.copy()
        REC.cur['hits'] = hits
        REC.cur['range_cells'] = r.copy()


# ------------------------------------------------------------------- CMap2D -------------
def flatten_contours(contours):
    return orc.flatten_contours(contours)


def render_contours_in_lidar(ranges, angles, flat_contours, lidar_xy):
    assert ranges.dtype == np.float32
    dirs = _libm_dirs(angles)
    before = ranges.copy()
    orc.render_contours(ranges, dirs, flat_contours, np.asarray(lidar_xy, np.float32))
    REC.log(fn='render_contours_in_lidar', ranges_in=before, angles=np.asarray(angles).copy(),
            flat=np.asarray(flat_contours, np.float32).copy(), lidar_xy=np.asarray(lidar_xy, np.float64).copy(),
            ranges_out=ranges.copy())
    REC.cur['segs'] = orc.contours_to_segments(flat_contours)


def _libm_dirs(angles):
    """beam directions through the SAME libm calls the C oracle uses."""
    import ctypes as C
    head = np.ascontiguousarray(np.asarray(angles).astype(np.float32))
    lin = head.astype(np.float64)  # heading_k = (float)(lin_k + 0.0) == head_k
    out_h = np.empty(len(head), np.float32)
    dirs = np.empty((len(head), 2), np.float32)
    orc.lib().nvo_beam_dirs(lin.ctypes.data_as(C.c_void_p), len(head), C.c_float(0.0),
                            out_h.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out_h, head)
    return dirs


class CSimAgent(object):
    def __init__(self, pose, state, vel):
        self.pose_2d_in_map_frame = np.asarray(pose, np.float32)
        self.state = np.asarray(state, np.float32)
        self.vel_in_map_frame = np.asarray(vel, np.float32)
        self.type = "legs"


class CMap2D(object):
    def __init__(self):
        self.resolution_ = 1.0

    def set_resolution(self, r):
        self.resolution_ = float(r)

    def render_agents_in_lidar(self, ranges, angles, agents, lidar_xy):
        assert ranges.dtype == np.float32
        discs = np.zeros((0, 3), np.float32)
        before = ranges.copy()
        if len(agents):
            discs = np.concatenate([orc.legs_to_discs(a.pose_2d_in_map_frame, a.state)
                                    for a in agents]).astype(np.float32)
            orc.render_discs(ranges, _libm_dirs(angles), discs, np.asarray(lidar_xy, np.float32))
        REC.log(fn='render_agents_in_lidar', ranges_in=before, angles=np.asarray(angles).copy(),
                poses=np.array([a.pose_2d_in_map_frame for a in agents], np.float32).reshape(-1, 3),
                states=np.array([a.state for a in agents], np.float32).reshape(-1, 3),
                vels=np.array([np.ravel(a.vel_in_map_frame) for a in agents], np.float32).reshape(len(agents), 2 if not len(agents) else -1),
                lidar_xy=np.asarray(lidar_xy, np.float64).copy(), ranges_out=ranges.copy())
        REC.finalize(ranges, discs)


# ------------------------------------------------------------------- pose2d -------------
def inverse_pose2d(p):
    x, y, th = p
    c, s = np.cos(-th), np.sin(-th)
    return np.array([c * (-x) - s * (-y), s * (-x) + c * (-y), -th])


def apply_tf_to_vel(vel, tf):
    th = tf[2]
    c, s = np.cos(th), np.sin(th)
    return np.array([c * vel[0] - s * vel[1], s * vel[0] + c * vel[1], vel[2]])


# ----------------------------------------------------------------- pyastar2d ------------
def astar_path(weights, start, goal, allow_diagonal=False):
    assert not allow_diagonal
    H, W = weights.shape
    start = (int(start[0]), int(start[1]))
    goal = (int(goal[0]), int(goal[1]))
    if not np.isfinite(weights[start]) or not np.isfinite(weights[goal]):
        return None
    g = {start: 0.0}
    parent = {}
    pq = [(abs(start[0] - goal[0]) + abs(start[1] - goal[1]), 0.0, start)]
    closed = set()
    while pq:
        _, gc, cur = heapq.heappop(pq)
        if cur in closed:
            continue
        closed.add(cur)
        if cur == goal:
            path = [cur]
            while cur in parent:
                cur = parent[cur]
                path.append(cur)
            return np.array(path[::-1], dtype=np.int64)
        for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            ni, nj = cur[0] + di, cur[1] + dj
            if ni < 0 or nj < 0 or ni >= H or nj >= W:
                continue
            w = weights[ni, nj]
            if not np.isfinite(w):
                continue
            ng = gc + float(w)
            if ng < g.get((ni, nj), np.inf):
                g[(ni, nj)] = ng
                parent[(ni, nj)] = cur
                h = (abs(ni - goal[0]) + abs(nj - goal[1])) * 1.0
                heapq.heappush(pq, (ng + h, ng, (ni, nj)))
    return None


# ------------------------------------------------------------------ install -------------
_installed = False


def install():
    global _installed
    if _installed:
        return
    from nav_gym_b200 import gym_shim
    gym_shim.install(force=True)

    m = types.ModuleType('range_libc')
    m.PyOMap, m.PyRayMarching = PyOMap, PyRayMarching
    sys.modules['range_libc'] = m
    m = types.ModuleType('CMap2D')
    m.flatten_contours, m.render_contours_in_lidar = flatten_contours, render_contours_in_lidar
    m.CMap2D, m.CSimAgent = CMap2D, CSimAgent
    sys.modules['CMap2D'] = m
    m = types.ModuleType('pose2d')
    m.inverse_pose2d, m.apply_tf_to_vel = inverse_pose2d, apply_tf_to_vel
    sys.modules['pose2d'] = m
    m = types.ModuleType('pyastar2d')
    m.astar_path = astar_path
    sys.modules['pyastar2d'] = m

    import torch
    _orig_load = torch.load

    def _load(path, *a, **kw):
        if str(path).endswith('human_policy.pth'):
            from nav_gym_env.human_policy import HumanPolicy
            gen = torch.random.get_rng_state()
            torch.manual_seed(1234)
            sd = HumanPolicy(frames=3, action_space=2).state_dict()
            torch.random.set_rng_state(gen)
            return sd
        return _orig_load(path, *a, **kw)

    torch.load = _load

    # Injected scan noise: the reference draws np.random.normal for the beams that are not at
    # range_max (env.py:438-440).  Draw from the same global stream but round each draw to a
    # float32-representable value, so that float32(range + noise) is ONE float32 addition —
    # the injected-noise tensor of the parity contract — and record it per beam.
    _orig_normal = np.random.normal

    def _normal(loc=0.0, scale=1.0, size=None):
        out = _orig_normal(loc, scale, size)
        if size is not None and REC.scans and REC.scans[-1]['noise'] is None:
            out = np.asarray(out).astype(np.float32).astype(np.float64)
            rec = REC.scans[-1]
            clipped = np.clip(rec['ranges_after'], 0, 25.0)
            mask = clipped != np.float32(25.0)
            if int(mask.sum()) == int(np.size(out)):
                full = np.zeros(len(clipped), np.float32)
                full[mask] = out.astype(np.float32)
                rec['noise'] = full
        return out

    np.random.normal = _normal
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    _installed = True


def make_env(**overrides):
    """gym.make('NavGym-v0') on the unmodified reference, kwargs as registered in
    nav_gym_env/__init__.py:6-38 with optional overrides."""
    install()
    import nav_gym_env  # noqa: F401  (runs the reference's register())
    import gym
    return gym.make('NavGym-v0', **overrides)
