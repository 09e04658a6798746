"""Mints tests/golden/*.npz by executing the reference's own env.py (unmodified, through
oracle/ref_harness.py) in the build container.  Run:  python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The fixtures pin the FIRST-PARTY arithmetic of the path
(kinematics keti_robot.py:64-93, cell mapping env.py:1228-1258, observation layout
env.py:443-462, reward / terminals / info env.py:464-589, crash rollback env.py:707-723,
step ordering env.py:591-728) to what the reference computes.  The three native calls inside
_compute_scan are served by the canonical stand-ins (range_libc / pymap2d sources are absent:
"parity unpinned" at that boundary), and what crossed that boundary — origin cell, headings,
hit cells, segments, discs, injected noise — is recorded per step.

Each trace is one episode: reset, then steps until done or max_steps.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), 'tests', 'golden')
NB = 512

TRACES = [
    # name, numpy seed, indoor_ratio, num_humans range, action mode, max_steps
    ('indoor_n1_random', 11, 1.0, [1, 1], 'random', 150),
    ('indoor_n8_random', 12, 1.0, [8, 8], 'random', 150),
    ('outdoor_n1_seek', 13, 0.0, [1, 1], 'seek', 400),
    ('outdoor_n12_seek', 14, 0.0, [12, 12], 'seek', 400),
    ('indoor_n5_seek', 15, 1.0, [5, 5], 'seek', 400),
    ('outdoor_n15_random', 16, 0.0, [15, 15], 'random', 150),
    # num_scan_stack = 3 (env.py:257-279): [pads | previous scans | current scan]
    ('indoor_n3_stack3', 17, 1.0, [3, 3], 'random', 150, 3),
    # min_turning_radius = 0.5 (env.py:595-600 clamps |v| >= |w| r and edits `action`, :725):
    # signed linear velocities reach both branches of the clamp
    ('indoor_n2_turnradius', 18, 1.0, [2, 2], 'random_signed', 150, 1, 0.5),
    ('outdoor_n6_turnradius_seek', 19, 0.0, [6, 6], 'seek', 400, 1, 0.5),
    # every constructor kwarg of the step path off its default (nav_gym_env/__init__.py:8-25): time
    # step, goal radius and all seven reward factors -- reward_forward_factor is 0 by default, so
    # only this trace sees the forward term of env.py:571
    ('indoor_n4_params', 20, 1.0, [4, 4], 'seek', 500, 1, 0.0,
     dict(time_step=0.1, distance_threshold=0.8, reward_scale=10., reward_success_factor=0.8,
          reward_crash_factor=1.5, reward_progress_factor=0.002, reward_forward_factor=0.02,
          reward_rotation_factor=0.01, reward_discomfort_factor=0.03)),
    # (num_humans = 0 is not a configuration of the reference: HumanPolicy.forward fails on the empty
    # batch, human_policy.py:45)
]


def _action(mode, env, rng):
    if mode == 'random':
        a = rng.uniform([0.0, -0.64], [0.5, 0.64])
    elif mode == 'random_signed':  # the reference does not clip actions (env.py:611-613)
        a = rng.uniform([-0.25, -0.64], [0.5, 0.64])
    else:  # head for the goal, with a little dither
        r = env.robot
        bearing = np.arctan2(r.gy - r.py, r.gx - r.px)
        err = np.arctan2(np.sin(bearing - r.theta), np.cos(bearing - r.theta))
        a = np.array([0.5 if abs(err) < 0.6 else 0.15, np.clip(1.5 * err, -0.64, 0.64)])
        a = a + rng.normal(0, 0.02, 2)
    # float32-representable actions, handed to the reference as float64
    return a.astype(np.float32).astype(np.float64)


def _robot_scans(crash):
    s = rh.REC.scans
    return (s[-2], s[-1]) if crash else (s[-1], None)


def _pack(recs, key, width):
    n = max([len(r[key]) for r in recs] + [1])
    out = np.zeros((len(recs), n, width), np.float32)
    cnt = np.zeros(len(recs), np.int32)
    for i, r in enumerate(recs):
        k = len(r[key])
        out[i, :k] = r[key]
        cnt[i] = k
    return out, cnt


def run_trace(name, seed, indoor_ratio, nh, mode, max_steps, stack=1, min_turning_radius=0.0, extra=None):
    np.random.seed(seed)
    import torch
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed + 1000)
    epr = dict(num_humans=(nh, 'int'), corridor_width=([3, 4], 'int'), iterations=([80, 150], 'int'),
               obstacle_number=([10, 10], 'int'), obstacle_width=([0.3, 1.0], 'float'),
               scan_noise_std=([0., 0.05], 'float'))
    env = rh.make_env(indoor_ratio=indoor_ratio, env_param_range=epr, num_scan_stack=stack,
                      min_turning_radius=min_turning_radius, **(extra or {}))
    NS = NB * stack
    rh.REC.clear()
    obs0 = env.reset()
    first, _ = _robot_scans(False)
    G = dict(
        map_data=env.map_info['data'].copy(), map_origin=np.array(env.map_info['origin'], np.float64),
        map_resolution=np.float64(env.map_info['resolution']),
        thr=env.scan_threshold.copy(), dthr=env.scan_discomfort_threshold.copy(),
        noise_std=np.float64(env.env_param['scan_noise_std']),
        start=np.array([env.robot.px, env.robot.py, env.robot.theta], np.float64),
        goal=np.array([env.robot.gx, env.robot.gy], np.float64),
        num_scan_stack=np.int32(stack), min_turning_radius=np.float64(min_turning_radius),
        obs0=obs0['observation'].astype(np.float64), hits0=first['hits'].astype(np.int16),
        cell0=first['ins'][0, :2].astype(np.int32),
        discs0=first['discs'], segs0=first['segs'],
        noise0=first['noise'] if first['noise'] is not None else np.zeros(NB, np.float32),
    )
    for k, v in (extra or {}).items():
        G['kw_' + k] = np.float64(v)
    recs1, recs2, rows = [], [], []
    for t in range(max_steps):
        a = _action(mode, env, rng)
        rh.REC.clear()
        obs, reward, done, info = env.step(a)
        crash = bool(info['is_crash'])
        s1, s2 = _robot_scans(crash)
        recs1.append(s1)
        recs2.append(s2 if s2 is not None else dict(discs=np.zeros((0, 3)), segs=np.zeros((0, 4)),
                                                    noise=None))
        rows.append(dict(
            action=a, scan=obs['observation'][NS - NB:NS].astype(np.float32), tail=obs['observation'][NS:],
            scan_stack=obs['observation'][:NS].astype(np.float32),
            achieved=obs['achieved_goal'], desired=obs['desired_goal'], reward=float(reward),
            done=bool(done), is_success=float(info['is_success']), is_crash=float(info['is_crash']),
            distance=float(info['distance']), hits=s1['hits'].astype(np.int16),
            cell=s1['ins'][0, :2].astype(np.int32),
            state=np.array([env.robot.px, env.robot.py, env.robot.theta]),
            steps=env.steps_since_reset))
        assert np.array_equal(obs['observation'][:NS].astype(np.float32).astype(np.float64),
                              obs['observation'][:NS])
        if done:
            break
    T = len(rows)
    G['actions'] = np.array([r['action'] for r in rows], np.float64)
    G['scan'] = np.array([r['scan'] for r in rows], np.float32)
    G['tail'] = np.array([r['tail'] for r in rows], np.float64)
    if stack > 1:
        G['scan_stack'] = np.array([r['scan_stack'] for r in rows], np.float32)
    G['achieved'] = np.array([r['achieved'] for r in rows], np.float64)
    G['desired'] = np.array([r['desired'] for r in rows], np.float64)
    G['reward'] = np.array([r['reward'] for r in rows], np.float64)
    G['done'] = np.array([r['done'] for r in rows], np.uint8)
    G['is_success'] = np.array([r['is_success'] for r in rows], np.uint8)
    G['is_crash'] = np.array([r['is_crash'] for r in rows], np.uint8)
    G['distance'] = np.array([r['distance'] for r in rows], np.float64)
    G['hits'] = np.array([r['hits'] for r in rows], np.int16)
    G['cell'] = np.array([r['cell'] for r in rows], np.int32)
    G['state'] = np.array([r['state'] for r in rows], np.float64)
    G['steps'] = np.array([r['steps'] for r in rows], np.int32)
    G['discs'], G['ndisc'] = _pack(recs1, 'discs', 3)
    G['segs'], G['nseg'] = _pack(recs1, 'segs', 4)
    noise = np.zeros((T, 2, NB), np.float32)
    for i in range(T):
        if recs1[i]['noise'] is not None:
            noise[i, 0] = recs1[i]['noise']
        if recs2[i].get('noise') is not None:
            noise[i, 1] = recs2[i]['noise']
    G['noise'] = noise
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **G)
    print('%-22s T=%3d done=%d success=%d crash=%d  ndisc<=%d nseg<=%d  %.1f KB' % (
        name, T, rows[-1]['done'], rows[-1]['is_success'], rows[-1]['is_crash'],
        G['ndisc'].max(), G['nseg'].max(), os.path.getsize(path) / 1024.0))


def known_answers():
    """KA vectors through the reference's own first-party functions."""
    rh.install()
    from nav_gym_env.keti_robot import KetiRobot
    from nav_gym_env import env as refenv
    rng = np.random.RandomState(7)
    n = 256
    st = np.column_stack([rng.uniform(0, 50, n), rng.uniform(0, 50, n), rng.uniform(-1, 7, n)])
    act = np.column_stack([rng.uniform(-0.2, 0.6, n), rng.uniform(-0.8, 0.8, n)])
    act = act.astype(np.float32).astype(np.float64)
    out = np.zeros((n, 3))
    for i in range(n):
        r = KetiRobot(st[i, 0], st[i, 1], st[i, 2], 0., 0., 0.2)
        r.set_vel(act[i, 0], act[i, 1])
        out[i] = (r.px, r.py, r.theta)
    mi = dict(resolution=0.05, origin=(0, 0), height=1000, width=1000)
    xy = rng.uniform(-1, 51, (4096, 2)).astype(np.float32)
    # values sitting on cell edges, where the float32/float64 rules can part
    edge = (np.arange(0, 1000, dtype=np.float64) * 0.05).astype(np.float32)
    xy = np.concatenate([xy, np.column_stack([edge, edge[::-1]])]).astype(np.float32)
    ij = np.array([refenv.xy_to_ij(p, mi) for p in xy], np.int64)  # NumPy-2 rule in this container
    np.savez_compressed(os.path.join(OUT, 'known_answers.npz'), kin_state=st, kin_action=act,
                        kin_out=out, cell_xy=xy, cell_ij=ij,
                        ka1=np.array([1.0975710555552805, 2.0242041665141244, 0.428]))
    r = KetiRobot(1., 2., 0.3, 0., 0., 0.2)
    r.set_vel(0.5, 0.64)
    print('KA1', repr(r.px), repr(r.py), repr(r.theta))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    known_answers()
    only = sys.argv[1:]
    for tr in TRACES:
        if only and tr[0] not in only:
            continue
        run_trace(*tr)
