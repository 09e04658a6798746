"""ctypes face of the CPU oracle (oracle/navgym_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of navgym_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the
product package nav_gym_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NAVGYM_ORACLE_VARIANT=_nofma selects the build whose march rounds x0 + dx * t in two steps
_SO = os.path.join(_HERE, '_build', 'libnavgym_oracle%s.so' % os.environ.get('NAVGYM_ORACLE_VARIANT', ''))

NB = 512
OBS_DIM = NB + 7
NS = 10
S_PX, S_PY, S_TH, S_GX, S_GY, S_PPX, S_PPY, S_PYAW, S_PV, S_PW = range(NS)
HIT_NONE = -32768


def build(force=False):
    src = os.path.join(_HERE, 'navgym_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE])
    return _SO


class MapT(C.Structure):
    _fields_ = [('W', C.c_int32), ('H', C.c_int32), ('offset', C.c_int64),
                ('ox', C.c_double), ('oy', C.c_double), ('res', C.c_double)]


class ParamsT(C.Structure):
    _fields_ = [('dt', C.c_double), ('dist_thresh', C.c_double), ('min_turn_radius', C.c_double),
                ('r_scale', C.c_double), ('r_success', C.c_double), ('r_crash', C.c_double),
                ('r_progress', C.c_double), ('r_forward', C.c_double), ('r_rotation', C.c_double),
                ('r_discomfort', C.c_double), ('range_max', C.c_float), ('t_stop', C.c_float),
                ('cell_rule', C.c_int32), ('max_disc', C.c_int32), ('max_seg', C.c_int32),
                ('num_scan_stack', C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.nvo_calc_range.restype = C.c_float
        _lib.nvo_xy_to_cell.restype = C.c_int32
        _lib.nvo_xy_to_cell.argtypes = [C.c_float, C.c_double, C.c_double, C.c_int, C.c_int]
        _lib.nvo_num_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


# ---- reference constants (keti_robot.py:12-48, __init__.py:6-38) -----------------------
ANGLE_MIN = -3.141592
ANGLE_MAX = 3.141592
ANGLE_INC = 0.0122718463
RANGE_MAX = 25.0
THRESHOLD_FOOTPRINT = [[0.6, 0.6], [-0.7, 0.6], [-0.7, -0.6], [0.6, -0.6]]
DISCOMFORT_FOOTPRINT = [[1.1, 1.1], [-0.7, 1.1], [-0.7, -1.1], [1.1, -1.1]]


def beam_table():
    """angles of env.py:388-390 before the heading is added (float64[512])."""
    return np.linspace(ANGLE_MIN, ANGLE_MAX - ANGLE_INC, NB)


def beam_dirs(theta32, lin=None):
    lin = beam_table() if lin is None else lin
    head = np.empty(NB, np.float32)
    dirs = np.empty((NB, 2), np.float32)
    lib().nvo_beam_dirs(_p(lin), NB, C.c_float(float(theta32)), _p(head), _p(dirs))
    return head, dirs


def edt(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    H, W = occ.shape
    out = np.empty((H, W), np.float32)
    lib().nvo_edt(_p(occ), H, W, _p(out))
    return out


def edt_sq(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    H, W = occ.shape
    out = np.empty((H, W), np.int32)
    lib().nvo_edt_sq(_p(occ), H, W, _p(out))
    return out


def grid_bfs(blocked, start_rc):
    """4-connected BFS distance (cells) over a blocked[H, W] grid, -1 = unreachable."""
    b = np.ascontiguousarray(blocked, np.uint8)
    out = np.empty(b.shape, np.int32)
    lib().nvo_grid_bfs(_p(b), b.shape[0], b.shape[1], int(start_rc[0]), int(start_rc[1]), _p(out))
    return out


def calc_range_many(dist, ins, max_range, t_stop=None, want_hits=False, want_steps=False):
    """range_libc PyRayMarching.calc_range_many restated; ranges in cells."""
    H, W = dist.shape
    ins = np.ascontiguousarray(ins, np.float32)
    n = ins.shape[0]
    outs = np.empty(n, np.float32)
    hits = np.empty((n, 2), np.int32) if want_hits else None
    steps = np.empty(n, np.int32) if want_steps else None
    t_stop = max_range if t_stop is None else t_stop
    lib().nvo_calc_range_many(_p(dist), W, H, _p(ins), _p(outs), n, C.c_float(max_range),
                              C.c_float(t_stop), _p(hits), _p(steps))
    res = [outs]
    if want_hits:
        res.append(hits)
    if want_steps:
        res.append(steps)
    return res[0] if len(res) == 1 else tuple(res)


def flatten_contours(contours):
    """pymap2d flatten_contours restated: list of polygons -> float32 [V,3] (id, x, y)."""
    rows = []
    for ci, c in enumerate(contours):
        for v in c:
            rows.append((float(ci), float(v[0]), float(v[1])))
    return np.asarray(rows, np.float32).reshape(-1, 3)


def render_contours(ranges, dirs, flat, lidar_xy):
    assert ranges.dtype == np.float32 and ranges.flags.c_contiguous
    flat = np.ascontiguousarray(flat, np.float32)
    dirs = np.ascontiguousarray(dirs, np.float32)
    lxy = np.ascontiguousarray(lidar_xy, np.float32)
    lib().nvo_render_contours(_p(ranges), _p(dirs), len(ranges), _p(flat), flat.shape[0], _p(lxy))


def render_segments(ranges, dirs, segs, lidar_xy):
    assert ranges.dtype == np.float32 and ranges.flags.c_contiguous
    segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 4)
    dirs = np.ascontiguousarray(dirs, np.float32)
    lxy = np.ascontiguousarray(lidar_xy, np.float32)
    lib().nvo_render_segments(_p(ranges), _p(dirs), len(ranges), _p(segs), segs.shape[0], _p(lxy))


def render_discs(ranges, dirs, discs, lidar_xy):
    assert ranges.dtype == np.float32 and ranges.flags.c_contiguous
    discs = np.ascontiguousarray(discs, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32)
    lxy = np.ascontiguousarray(lidar_xy, np.float32)
    lib().nvo_render_discs(_p(ranges), _p(dirs), len(ranges), _p(discs), discs.shape[0], _p(lxy))


def contours_to_segments(flat):
    """closed polygons (flat [V,3]) -> [S,4] segments, the closing edge included."""
    segs = []
    flat = np.asarray(flat, np.float32).reshape(-1, 3)
    ids = flat[:, 0]
    s = 0
    V = len(flat)
    while s < V:
        e = s
        while e + 1 < V and ids[e + 1] == ids[s]:
            e += 1
        for v in range(s, e + 1):
            w = s if v == e else v + 1
            segs.append((flat[v, 1], flat[v, 2], flat[w, 1], flat[w, 2]))
        s = e + 1
    return np.asarray(segs, np.float32).reshape(-1, 4)


def footprint_threshold(footprint):
    """_make_scan_threshold / _make_scan_discomfort_threshold (env.py:162-180): the scan of
    the footprint polygon rendered from the origin at heading 0, clipped to [0, 25]."""
    ranges = np.full(NB, RANGE_MAX, np.float32)
    _, dirs = beam_dirs(np.float32(0.0))
    render_contours(ranges, dirs, flatten_contours([footprint]), np.zeros(2, np.float32))
    return np.clip(ranges, 0, RANGE_MAX)


def legs_to_discs(pose, dist_travelled):
    """pymap2d CSimAgent 'legs' model restated (constants from memory of
    the upstream package — PARITY UNPINNED): two leg discs of radius 0.03 m whose fore-aft /
    lateral offsets oscillate with the distance travelled in the base frame."""
    x, y, th = [float(v) for v in pose]
    s = [float(v) for v in dist_travelled]
    leg_radius, side_off, side_amp, front_amp = 0.03, 0.1, 0.1, 0.3
    front = front_amp * np.cos(s[0] * 2.0 / front_amp + s[2])
    side = side_amp * np.cos(s[1] * 2.0 / side_amp + s[2])
    out = np.empty((2, 3), np.float64)
    for i, (lx, ly) in enumerate(((front, side + side_off), (-front, -side - side_off))):
        out[i, 0] = x + np.cos(th) * lx - np.sin(th) * ly
        out[i, 1] = y + np.sin(th) * lx + np.cos(th) * ly
        out[i, 2] = leg_radius
    return out.astype(np.float32)


def default_params(**kw):
    p = dict(dt=0.2, dist_thresh=0.5, min_turn_radius=0.0, r_scale=15.0, r_success=1.0,
             r_crash=1.0, r_progress=0.001, r_forward=0.0, r_rotation=0.005, r_discomfort=0.01,
             range_max=RANGE_MAX, t_stop=1e12, cell_rule=0, max_disc=0, max_seg=0, num_scan_stack=1)
    p.update(kw)
    return p


class OracleBatch(object):
    """B environments stepped in lockstep by the C oracle (nvo_step_batch)."""

    def __init__(self, maps, map_id, start, goal, theta, params=None, max_disc=0, max_seg=0):
        """maps: list of dicts like the reference's map_info (data int8 [H,W], origin,
        resolution, width, height); map_id int[B]; start/goal float64 [B,2]; theta [B]."""
        self.params = default_params(**(params or {}))
        self.params['max_disc'] = max_disc
        self.params['max_seg'] = max_seg
        self.B = B = len(map_id)
        self.map_id = np.ascontiguousarray(map_id, np.int32)
        edts, off = [], 0
        self.maps_c = (MapT * len(maps))()
        for i, m in enumerate(maps):
            d = edt(np.asarray(m['data']) >= 0.1)  # env.py:339
            edts.append(d.ravel())
            self.maps_c[i] = MapT(int(m['width']), int(m['height']), off, float(m['origin'][0]),
                                  float(m['origin'][1]), float(m['resolution']))
            off += d.size
        self.edt_pool = np.concatenate(edts)
        self.lin = beam_table()
        self.thr = footprint_threshold(THRESHOLD_FOOTPRINT)
        self.dthr = footprint_threshold(DISCOMFORT_FOOTPRINT)
        self.state = np.zeros((NS, B), np.float64)
        self.state[S_PX], self.state[S_PY] = np.asarray(start, np.float64).T
        self.state[S_TH] = theta
        self.state[S_GX], self.state[S_GY] = np.asarray(goal, np.float64).T
        self.steps = np.zeros(B, np.int32)
        S = self.S = max(int(self.params['num_scan_stack']), 1)
        self.obs = np.zeros((B, S * NB + 7), np.float32)
        self.hist = np.zeros((B, max(S - 1, 1), NB), np.float32)
        self.nhist = np.zeros(B, np.int32)
        self.tail64 = np.zeros((B, 7), np.float64)
        self.reward = np.zeros(B, np.float64)
        self.done = np.zeros(B, np.uint8)
        self.is_success = np.zeros(B, np.uint8)
        self.is_crash = np.zeros(B, np.uint8)
        self.distance = np.zeros(B, np.float64)
        self.hits = np.zeros((B, NB, 2), np.int16)

    def _cparams(self):
        p = self.params
        return ParamsT(p['dt'], p['dist_thresh'], p['min_turn_radius'], p['r_scale'], p['r_success'],
                       p['r_crash'], p['r_progress'], p['r_forward'], p['r_rotation'],
                       p['r_discomfort'], p['range_max'], p['t_stop'], p['cell_rule'],
                       p['max_disc'], p['max_seg'], p['num_scan_stack'])

    def _geom(self, discs, ndisc, segs, nseg):
        md, ms = self.params['max_disc'], self.params['max_seg']
        if discs is not None:
            discs = np.ascontiguousarray(discs, np.float32).reshape(self.B, md, 3)
            ndisc = np.ascontiguousarray(ndisc, np.int32)
        if segs is not None:
            segs = np.ascontiguousarray(segs, np.float32).reshape(self.B, ms, 4)
            nseg = np.ascontiguousarray(nseg, np.int32)
        return discs, ndisc, segs, nseg

    def reset_obs(self, discs=None, ndisc=None, segs=None, nseg=None, noise=None, want_hits=True):
        discs, ndisc, segs, nseg = self._geom(discs, ndisc, segs, nseg)
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float32).reshape(self.B, 2, NB)
        cp = self._cparams()
        lib().nvo_reset_obs_batch(C.byref(cp), self.B, self.maps_c, _p(self.edt_pool),
                                  _p(self.map_id), _p(self.lin), _p(self.state), _p(self.steps),
                                  _p(discs), _p(ndisc), _p(segs), _p(nseg), _p(noise),
                                  _p(self.obs), _p(self.tail64),
                                  _p(self.hits) if want_hits else None,
                                  _p(self.hist) if self.S > 1 else None, _p(self.nhist) if self.S > 1 else None)
        return self.obs

    def step(self, actions, discs=None, ndisc=None, segs=None, nseg=None, noise=None,
             want_hits=True):
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.B, 2)
        discs, ndisc, segs, nseg = self._geom(discs, ndisc, segs, nseg)
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float32).reshape(self.B, 2, NB)
        cp = self._cparams()
        lib().nvo_step_batch(C.byref(cp), self.B, self.maps_c, _p(self.edt_pool), _p(self.map_id),
                             _p(self.lin), _p(self.thr), _p(self.dthr), _p(self.state),
                             _p(self.steps), _p(actions), _p(discs), _p(ndisc), _p(segs), _p(nseg),
                             _p(noise), _p(self.obs), _p(self.tail64), _p(self.reward),
                             _p(self.done), _p(self.is_success), _p(self.is_crash),
                             _p(self.distance), _p(self.hits) if want_hits else None,
                             _p(self.hist) if self.S > 1 else None, _p(self.nhist) if self.S > 1 else None)
        return self.obs, self.reward, self.done

    def set_threads(self, n):
        os.environ['OMP_NUM_THREADS'] = str(n)


def num_threads():
    return lib().nvo_num_threads()


def use_all_cores():
    """Run the OpenMP loops on every host core (launchers such as torchrun export
    OMP_NUM_THREADS=1, which would otherwise throttle the CPU baseline)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    lib().nvo_set_num_threads(int(n))
    return num_threads()
