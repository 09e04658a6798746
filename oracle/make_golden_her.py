"""Known answers for the HER entry points: the synthetic batch of tests/synth.py her_batch()
pushed through the UNMODIFIED reference's compute_rewards / compute_terminals (env.py:491-589),
as float64 observations (what env.step returns) -> tests/golden/her_batch.npz (outputs only; the
batch is regenerated from its seed).  Test infrastructure; run in the build container:
    python oracle/make_golden_her.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_harness as rh  # noqa: E402
import synth  # noqa: E402

if __name__ == '__main__':
    np.random.seed(0)
    env = rh.make_env().unwrapped
    b = synth.her_batch()
    rows = synth.her_rows(b, env.scan_threshold, env.scan_discomfort_threshold)
    obs = dict(observation=rows.astype(np.float64), desired_goal=b['goal'].astype(np.float64),
               achieved_goal=rows[:, 514:516].astype(np.float64))
    rew = env.compute_rewards(None, obs)
    done = env.compute_terminals(obs)
    out = os.path.join(ROOT, 'tests', 'golden', 'her_batch.npz')
    np.savez_compressed(out, reward=np.asarray(rew, np.float64), done=np.asarray(done, bool),
                        thr=np.asarray(env.scan_threshold, np.float32),
                        dthr=np.asarray(env.scan_discomfort_threshold, np.float32))
    k = b['kind']
    print('rows %d: clear %d, discomfort %d, crash %d, at goal %d; done %d; reward range [%.3f, %.3f]; %.1f KB' % (
        len(k), (k == 0).sum(), (k == 1).sum(), (k == 2).sum(), (k == 3).sum(), int(np.sum(done)),
        rew.min(), rew.max(), os.path.getsize(out) / 1024.0))
    # every discomfort row must score a discomfort penalty in (-0.15, 0]: the band was hit
    disc = k == 1
    base = 0.015 * (np.linalg.norm(b['goal'] - rows[:, 512:514], axis=1) - np.linalg.norm(b['goal'] - rows[:, 514:516], axis=1)) \
        - 0.075 * rows[:, 517].astype(np.float64) ** 2
    pen = rew[disc] - base[disc]
    at_goal = np.linalg.norm(b['goal'] - rows[:, 514:516], axis=1) < 0.5
    print('discomfort penalties in [%.4f, %.4f] (rows not at the goal)' % (pen[~at_goal[disc]].min(), pen[~at_goal[disc]].max()))
