"""Mints tests/golden/humans_*.npz: what the reference's pedestrian pipeline computes per step
(env.py:617-693), recorded while the UNMODIFIED env.py runs through oracle/ref_harness.py.
Run:  python oracle/make_golden_humans.py

TEST INFRASTRUCTURE ONLY.  human_policy.pth is missing from the checkout, so the policy runs
with the harness's seeded random-init weights (torch.manual_seed(1234); HumanPolicy(3, 2)); the
tests rebuild the same weights from that seed.  Per step the trace holds, for every pedestrian:
the policy inputs (newest scan, local goal, previous action), the policy mean, the pose before
and after Human.set_vel, the leg-gait odometry, the scan it takes afterwards, and the robot
pose those scans saw.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), 'tests', 'golden')
TRACES = [('humans_indoor_n6', 21, 1.0, 6, 40), ('humans_outdoor_n10', 22, 0.0, 10, 30)]


def run(name, seed, indoor_ratio, nh, steps):
    import torch
    np.random.seed(seed)
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed + 1000)
    epr = dict(num_humans=([nh, nh], 'int'), corridor_width=([3, 4], 'int'), iterations=([80, 150], 'int'),
               obstacle_number=([10, 10], 'int'), obstacle_width=([0.3, 1.0], 'float'),
               scan_noise_std=([0., 0.05], 'float'))
    env = rh.make_env(indoor_ratio=indoor_ratio, env_param_range=epr)
    env.reset()
    calls = []
    policy = env.human_policy
    orig_forward = policy.forward

    def forward(x, goal, speed):
        out = orig_forward(x, goal, speed)
        calls.append((x.numpy().copy(), goal.numpy().copy(), speed.numpy().copy(), out[3].detach().numpy().copy()))
        return out
    policy.forward = forward
    moved = {}
    orig_set_vel = env.robot.set_vel

    def set_vel(v, w):
        orig_set_vel(v, w)
        moved['pose'] = (env.robot.px, env.robot.py, env.robot.theta)
    env.robot.set_vel = set_vel

    def poses():
        return np.array([[h.px, h.py, h.theta] for h in env.humans], np.float64)

    def scans():
        return np.array([q[-1]['observation'][2 * 512:3 * 512] for q in env.prev_humans_obs_queue], np.float32)
    G = dict(map_data=env.map_info['data'].copy(), map_origin=np.array(env.map_info['origin'], np.float64),
             map_resolution=np.float64(env.map_info['resolution']),
             v_pref=np.array([h.v_pref for h in env.humans], np.float64),
             has_legs=np.array([h.has_legs for h in env.humans], np.uint8),
             pose0=poses(), scan0=scans(),
             robot0=np.array([env.robot.px, env.robot.py, env.robot.theta], np.float64))
    rows = []
    for t in range(steps):
        a = rng.uniform([0.0, -0.64], [0.5, 0.64]).astype(np.float32).astype(np.float64)
        before = poses()
        rh.REC.clear()
        obs, reward, done, info = env.step(a)
        hrecs = rh.REC.scans[:nh]  # env.py:683-693 scans the pedestrians first, in order
        S = max(len(r['segs']) for r in hrecs)
        segs = np.zeros((nh, S, 4), np.float32)
        nseg = np.zeros(nh, np.int32)
        for i, r in enumerate(hrecs):
            segs[i, :len(r['segs'])] = r['segs']
            nseg[i] = len(r['segs'])
        x, goal, speed, mean = calls[-1]
        assert np.array_equal(x[:, 0], x[:, 2])  # env.py:647 hands the newest scan in all frames
        rows.append(dict(
            pose_before=before, scan_in=x[:, 2].copy(), goal_local=goal, speed=speed, mean=mean,
            goal_world=np.array([[h.gx, h.gy] for h in env.humans], np.float64),
            pose_after=poses(), vel=np.array([[h.vx, h.vy] for h in env.humans], np.float64),
            dist=env.distances_travelled_in_base_frame.copy(), scan_out=scans(),
            robot_moved=np.array(moved['pose'], np.float64), crash=np.uint8(info['is_crash']),
            segs_out=segs, nseg_out=nseg))
        if done:
            break
    for k in rows[0]:
        G[k] = np.array([r[k] for r in rows])
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **G)
    print('%-20s T=%d humans=%d  %.1f KB' % (name, len(rows), nh, os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    for tr in TRACES:
        run(*tr)
