"""Replays a golden trace (tests/golden/*.npz, minted from the reference's own env.py by
oracle/make_golden.py) through any lockstep stepper and compares every step."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NB = 512


def trace_names():
    """the robot-step traces (humans_*.npz hold the pedestrian pipeline, see human_trace_names)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz'))
                  if os.path.basename(p) not in ('known_answers.npz', 'bench_world.npz', 'native_calls.npz', 'her_batch.npz')
                  and not os.path.basename(p).startswith('humans_'))


def human_trace_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'humans_*.npz')))


def human_env_segments(G, t):
    """Step t of a humans_* trace -> (segs [S, 4], skip [P, 2]): the environment's segment list
    [robot | pedestrian 0 | pedestrian 1 ...] (5 per closed footprint, as the reference's
    contours carry them) and, per pedestrian, its own slice to leave out.  Rebuilt from the
    per-pedestrian lists the reference handed to render_contours_in_lidar
    ([robot] + the other pedestrians, env.py:683-688)."""
    P = G['segs_out'].shape[1]
    polys = [G['segs_out'][t, 0, :5]]
    for j in range(P):
        src = 1 if j == 0 else 0          # a list that contains pedestrian j
        k = j if j > src else j + 1       # its position among that list's pedestrians, 1-based
        polys.append(G['segs_out'][t, src, 5 * k:5 * k + 5])
    segs = np.concatenate(polys).astype(np.float32)
    skip = np.array([[5 * (1 + j), 5] for j in range(P)], np.int32)
    return segs, skip


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def map_info(G):
    d = G['map_data']
    return dict(data=d, origin=tuple(G['map_origin']), resolution=float(G['map_resolution']),
                width=d.shape[1], height=d.shape[0])


def env_kwargs(G):
    """Constructor kwargs a trace was minted with off their defaults (reference names:
    time_step, distance_threshold, reward_*), stored as kw_<name> scalars."""
    return {k[3:]: float(G[k]) for k in G if k.startswith('kw_')}


def geom_dims(G):
    md = max(G['discs'].shape[1], len(G['discs0']), 1)
    ms = max(G['segs'].shape[1], len(G['segs0']), 1)
    return md, ms


def pad(a, n, w):
    out = np.zeros((1, n, w), np.float32)
    a = np.asarray(a, np.float32).reshape(-1, w)
    out[0, :len(a)] = a
    return out, np.array([len(a)], np.int32)


def replay(G, stepper, pose_tol=1e-9, reward_tol=1e-9, check_hits=True):
    """stepper: object with reset_obs(...)/step(...) and numpy attrs obs, tail64, reward,
    done, is_success, is_crash, distance, hits, state (rows px,py,th first), steps.
    Integer/flag outputs and the float32 scan are compared bit-exactly."""
    md, ms = geom_dims(G)
    S = int(G['num_scan_stack']) if 'num_scan_stack' in G else 1
    NS = S * NB
    d0, nd0 = pad(G['discs0'], md, 3)
    s0, ns0 = pad(G['segs0'], ms, 4)
    n0 = np.zeros((1, 2, NB), np.float32)
    n0[0, 0] = G['noise0']
    stepper.reset_obs(d0, nd0, s0, ns0, n0)
    obs = np.asarray(stepper.obs)[0]
    assert np.array_equal(obs[:NS], G['obs0'][:NS].astype(np.float32)), 'first scan (stack)'
    assert np.allclose(np.asarray(stepper.tail64)[0], G['obs0'][NS:], rtol=0, atol=pose_tol)
    if check_hits:
        assert np.array_equal(_mask_hits(np.asarray(stepper.hits)[0], stepper), _mask_hits(G['hits0'], stepper))
    T = len(G['actions'])
    for t in range(T):
        dd, nd = pad(G['discs'][t][:G['ndisc'][t]], md, 3)
        ss, ns = pad(G['segs'][t][:G['nseg'][t]], ms, 4)
        stepper.step(G['actions'][t][None].astype(np.float32), dd, nd, ss, ns, G['noise'][t][None])
        obs = np.asarray(stepper.obs)[0]
        tag = 'step %d' % t
        assert np.array_equal(obs[NS - NB:NS], G['scan'][t]), tag + ' scan'
        if S > 1:
            assert np.array_equal(obs[:NS], G['scan_stack'][t]), tag + ' scan stack'
        assert np.allclose(np.asarray(stepper.tail64)[0], G['tail'][t], rtol=0, atol=pose_tol), tag
        assert np.allclose(obs[NS:], G['tail'][t].astype(np.float32), rtol=1e-6, atol=1e-6), tag
        assert abs(float(np.asarray(stepper.reward)[0]) - G['reward'][t]) <= reward_tol, tag + ' reward'
        assert int(np.asarray(stepper.done)[0]) == int(G['done'][t]), tag + ' done'
        assert int(np.asarray(stepper.is_success)[0]) == int(G['is_success'][t]), tag
        assert int(np.asarray(stepper.is_crash)[0]) == int(G['is_crash'][t]), tag
        assert abs(float(np.asarray(stepper.distance)[0]) - G['distance'][t]) <= 1e-6, tag
        assert int(np.asarray(stepper.steps)[0]) == int(G['steps'][t]), tag + ' step count'
        st = np.asarray(stepper.state)[:3, 0]
        assert np.allclose(st, G['state'][t], rtol=0, atol=pose_tol), tag + ' pose'
        if check_hits:
            assert np.array_equal(_mask_hits(np.asarray(stepper.hits)[0], stepper), _mask_hits(G['hits'][t], stepper)), \
                tag + ' hit cells'
    return T


def far_filter(h):
    """Hit cells farther than 500.5 cells (25 m) from the origin are not observable in the
    reference (env.py:435 clips the range): blank them on both sides of the comparison."""
    h = np.array(h, np.int16)
    d2 = h[:, 0].astype(np.int64) ** 2 + h[:, 1].astype(np.int64) ** 2
    h[d2 > 500.5 ** 2] = -32768
    return h


def _mask_hits(h, stepper):
    """A stepper that stops marching at 25 m (+2 cells) reports 'no hit' for farther cells;
    the reference marches on to W*H cells (env.py:337) and clips the range afterwards."""
    return far_filter(h) if getattr(stepper, 'early_stop', False) else np.asarray(h)
