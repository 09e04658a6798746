"""GPU-backed stand-ins with the Python surface of the two native packages the reference calls
on its hot path (SURVEY §8b "inner native boundary"):

  range_libc :  PyOMap(bool[H,W]), PyRayMarching(omap, max_range).calc_range_many(ins, outs)
                (reference call sites env.py:337-340, 420-426)
  CMap2D     :  flatten_contours, render_contours_in_lidar, CMap2D().render_agents_in_lidar,
                CSimAgent   (reference call sites env.py:14, 102-103, 398-402, 428-432)

Same names, argument meaning, in-place output and error behaviour (dtype / contiguity
mismatches raise ValueError like Cython typed buffers do); the arithmetic runs in the sm_100a
kernels behind include/navgym_b200.h.  :func:`install_as_reference_natives` registers them in
``sys.modules`` under the reference's import names, which is the binding a maintainer of the
reference would use.
"""
import ctypes as C
import sys
import types

import numpy as np

from . import _lib
from .robot import legs_to_discs


def _need(arr, dtype, ndim, name):
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for %s: expected %s" % (name, np.dtype(dtype).name))
    if arr.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions for %s (expected %d, got %d)"
                         % (name, ndim, arr.ndim))
    if not arr.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous: %s" % name)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class PyOMap(object):
    """range_libc.PyOMap(np.ndarray[bool, ndim=2]) — occupancy grid, True = occupied."""

    def __init__(self, arr):
        if not isinstance(arr, np.ndarray) or arr.ndim != 2:
            raise ValueError("PyOMap expects a 2-D ndarray")
        self.grid = np.ascontiguousarray(arr.astype(np.uint8))
        self.height, self.width = self.grid.shape


class PyRayMarching(object):
    """range_libc.PyRayMarching(omap, max_range): builds the exact EDT on the device and keeps
    it for the lifetime of the object (one per episode in the reference, env.py:338-340)."""

    def __init__(self, omap, max_range):
        self._lib = _lib.require_device()
        self.max_range = float(max_range)
        self.height, self.width = omap.height, omap.width
        self._h = self._lib.navgym_raymarching_create_host(_vp(omap.grid), omap.height, omap.width,
                                                           C.c_float(self.max_range))
        if not self._h:
            raise RuntimeError('navgym_raymarching_create_host failed')

    def calc_range_many(self, ins, outs):
        """ins float32[N,3] = (x_cell, y_cell, heading_rad); outs float32[N] written in place,
        ranges in cells."""
        _need(ins, np.float32, 2, 'ins')
        _need(outs, np.float32, 1, 'outs')
        if ins.shape[1] != 3 or ins.shape[0] != outs.shape[0]:
            raise ValueError('ins must be [N,3] and outs [N]')
        _lib.check(self._lib.navgym_raymarching_calc_range_many_host(
            C.c_void_p(self._h), _vp(ins), _vp(outs), ins.shape[0]), 'calc_range_many')

    def calc_range_many_with_hits(self, ins):
        """Extension: also returns the hit cell minus the origin cell (int16 [N,2])."""
        import torch
        _need(ins, np.float32, 2, 'ins')
        n = ins.shape[0]
        d_ins = torch.from_numpy(ins).cuda()
        d_out = torch.empty(n, dtype=torch.float32, device='cuda')
        d_hit = torch.empty(n, 2, dtype=torch.int16, device='cuda')
        edt = self._lib.navgym_raymarching_edt_dev(C.c_void_p(self._h))
        _lib.check(self._lib.navgym_calc_range_many(
            C.c_void_p(edt), self.width, self.height, C.c_void_p(d_ins.data_ptr()),
            C.c_void_p(d_out.data_ptr()), n, C.c_float(self.max_range), C.c_float(self.max_range),
            C.c_void_p(d_hit.data_ptr()),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'calc_range_many')
        return d_out.cpu().numpy(), d_hit.cpu().numpy()

    def edt_host(self):
        """The distance image (float32 [H,W], cells) copied back to the host."""
        out = np.empty((self.height, self.width), np.float32)
        _lib.check(self._lib.navgym_raymarching_edt_host(C.c_void_p(self._h), _vp(out)), 'edt_host')
        return out

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._lib.navgym_raymarching_destroy(C.c_void_p(h))


def flatten_contours(contours):
    """list of polygons (lists of [x, y]) -> float32 [V,3] rows (contour id, x, y)."""
    rows = []
    for ci, c in enumerate(contours):
        for v in c:
            rows.append((float(ci), float(v[0]), float(v[1])))
    return np.asarray(rows, np.float32).reshape(-1, 3)


def _closed_segments(flat):
    flat = np.asarray(flat, np.float32).reshape(-1, 3)
    segs = []
    s, V = 0, len(flat)
    while s < V:
        e = s
        while e + 1 < V and flat[e + 1, 0] == flat[s, 0]:
            e += 1
        for v in range(s, e + 1):
            w = s if v == e else v + 1
            segs.append((flat[v, 1], flat[v, 2], flat[w, 1], flat[w, 2]))
        s = e + 1
    return np.ascontiguousarray(np.asarray(segs, np.float32).reshape(-1, 4))


def _render(ranges, angles, segs, discs, lidar_xy):
    _need(ranges, np.float32, 1, 'ranges')
    head = np.ascontiguousarray(np.asarray(angles).astype(np.float32))
    if head.shape[0] != ranges.shape[0]:
        raise ValueError('ranges and angles differ in length')
    lib = _lib.require_device()
    S = 0 if segs is None else len(segs)
    D = 0 if discs is None else len(discs)
    if S + D == 0:
        return
    _lib.check(lib.navgym_render_in_lidar_host(
        _vp(ranges), _vp(head), len(head), _vp(segs) if S else None, S,
        _vp(discs) if D else None, D, C.c_float(float(lidar_xy[0])), C.c_float(float(lidar_xy[1]))),
        'render_in_lidar_host')


def render_contours_in_lidar(ranges, angles, flat_contours, lidar_xy):
    """In place: ranges[k] = min(ranges[k], distance along beam k to any polygon edge)."""
    _render(ranges, angles, _closed_segments(flat_contours), None, lidar_xy)


class CSimAgent(object):
    def __init__(self, pose, state, vel):
        self.pose_2d_in_map_frame = np.asarray(pose, np.float32)
        self.state = np.asarray(state, np.float32)
        self.vel_in_map_frame = np.asarray(vel, np.float32)
        self.type = "legs"
        self.leg_radius = 0.03


class CMap2D(object):
    def __init__(self):
        self.resolution_ = 1.0
        self.origin = np.zeros(2, np.float32)

    def set_resolution(self, r):
        self.resolution_ = float(r)

    def resolution(self):
        return self.resolution_

    def render_agents_in_lidar(self, ranges, angles, agents, lidar_xy):
        """In place: min-merge the two leg discs of every agent into the scan."""
        if not len(agents):
            return
        for a in agents:
            if a.type != "legs":
                raise NotImplementedError
        discs = np.concatenate([legs_to_discs(a.pose_2d_in_map_frame, a.state) for a in agents])
        _render(ranges, angles, None, np.ascontiguousarray(discs, np.float32), lidar_xy)


# ---- pose2d / pyastar2d: host-side helpers of the reset path and the leg odometry ----------
def inverse_pose2d(pose):
    """pose2d.inverse_pose2d (call site env.py:252): the transform that undoes (x, y, theta)."""
    x, y, th = (float(v) for v in pose)
    c, s = np.cos(th), np.sin(th)
    return np.array([-(c * x + s * y), s * x - c * y, -th])


def apply_tf_to_vel(vel, tf):
    """pose2d.apply_tf_to_vel (call site env.py:254): rotate (vx, vy) by the transform's angle,
    the rotation rate is frame-independent."""
    c, s = np.cos(tf[2]), np.sin(tf[2])
    return np.array([c * vel[0] - s * vel[1], s * vel[0] + c * vel[1], vel[2]])


def astar_path(weights, start, goal, allow_diagonal=False):
    """pyastar2d.astar_path on the reference's planning grid (env.py:343-351: inf where blocked,
    one uniform cost elsewhere, 4-connected): with uniform costs every 4-connected shortest path
    is optimal, so one is read off the library's BFS distance field from the goal.  Returns int
    [L, 2] cells from start to goal inclusive, or None when there is no path."""
    if allow_diagonal:
        raise NotImplementedError('the reference plans 4-connected (env.py:351)')
    w = np.asarray(weights)
    si, sj, gi, gj = int(start[0]), int(start[1]), int(goal[0]), int(goal[1])
    blocked = ~np.isfinite(w)
    if blocked[si, sj] or blocked[gi, gj]:
        return None
    from .maps import grid_bfs
    d = grid_bfs(blocked, (gi, gj))
    if d[si, sj] < 0:
        return None
    path = [(si, sj)]
    i, j = si, sj
    H, W = d.shape
    while (i, j) != (gi, gj):
        for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            ni, nj = i + di, j + dj
            if 0 <= ni < H and 0 <= nj < W and d[ni, nj] == d[i, j] - 1:
                i, j = ni, nj
                break
        path.append((i, j))
    return np.array(path, dtype=np.int64)


def install_as_reference_natives():
    """Expose these classes as ``range_libc`` and ``CMap2D`` so the reference's env.py picks
    them up unchanged; ``pose2d`` and ``pyastar2d`` (the other two un-vendored imports of
    env.py:13-17) are registered too when the real packages are not installed."""
    import importlib.util
    for name, attrs in (('pose2d', dict(inverse_pose2d=inverse_pose2d, apply_tf_to_vel=apply_tf_to_vel)),
                        ('pyastar2d', dict(astar_path=astar_path))):
        if name not in sys.modules and importlib.util.find_spec(name) is None:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    m = types.ModuleType('range_libc')
    m.PyOMap, m.PyRayMarching = PyOMap, PyRayMarching
    sys.modules['range_libc'] = m
    m = types.ModuleType('CMap2D')
    m.flatten_contours, m.render_contours_in_lidar = flatten_contours, render_contours_in_lidar
    m.CMap2D, m.CSimAgent = CMap2D, CSimAgent
    sys.modules['CMap2D'] = m
