"""Mints tests/golden/native_calls.npz: the transcript of every call the UNMODIFIED reference
env.py makes across its native boundary (range_libc / pymap2d, env.py:337-340, 425-432) during
one short episode -- arguments exactly as the reference passed them, results as the oracle-backed
stand-ins of oracle/ref_harness.py returned them.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  The GPU test
tests/test_gpu_parity.py::test_level1_native_call_transcript replays the transcript through
nav_gym_b200.natives -- the product's level-1 binding -- and demands
the same results bit for bit: the reference's own call sequence, executed on the device.
The native semantics themselves remain "parity unpinned" (third-party sources absent).

    python oracle/make_golden_native_calls.py
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'native_calls.npz')


def main(seed=31, steps=6, nh=4):
    np.random.seed(seed)
    import torch
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed + 1000)
    epr = dict(num_humans=([nh, nh], 'int'), corridor_width=([3, 4], 'int'), iterations=([80, 150], 'int'),
               obstacle_number=([10, 10], 'int'), obstacle_width=([0.3, 1.0], 'float'),
               scan_noise_std=([0., 0.05], 'float'))
    env = rh.make_env(indoor_ratio=0.0, env_param_range=epr)   # outdoor: 400 x 400 cells keep the fixture small
    rh.REC.clear()
    rh.REC.calls = []
    env.reset()
    for t in range(steps):
        a = rng.uniform([0.0, -0.64], [0.5, 0.64]).astype(np.float32).astype(np.float64)
        _, _, done, _ = env.step(a)
        if done:
            break
    calls, rh.REC.calls = rh.REC.calls, None
    kinds = ['PyRayMarching', 'calc_range_many', 'render_contours_in_lidar', 'render_agents_in_lidar']
    G = {'kind': np.array([kinds.index(c['fn']) for c in calls], np.int8), 'kind_names': np.array(kinds)}
    n_legs = n_box = 0
    for i, c in enumerate(calls):
        for k, v in c.items():
            if k == 'fn':
                continue
            v = np.asarray(v)
            if k == 'occ':
                G['c%d_occ_bits' % i] = np.packbits(v.astype(np.uint8).reshape(-1))
                G['c%d_occ_shape' % i] = np.array(v.shape, np.int32)
            else:
                G['c%d_%s' % (i, k)] = v
        n_legs += c['fn'] == 'render_agents_in_lidar' and len(c['poses']) > 0
        n_box += c['fn'] == 'render_contours_in_lidar'
    np.savez_compressed(OUT, **G)
    cnt = {k: int((G['kind'] == i).sum()) for i, k in enumerate(kinds)}
    print('%s: %d calls %s (agent renders with legs: %d), %.0f KB' % (OUT, len(calls), cnt, n_legs, os.path.getsize(OUT) / 1024))


if __name__ == '__main__':
    main()
