"""The reference's OWN Python step loop, timed: nav_gym_env/env.py executed
unmodified through oracle/ref_harness.make_env() (functional stand-ins only for the five absent
pip packages; the lidar natives are the compiled C checker, so the loop is not handicapped by a
Python raycast), stepped as the reference's smoke loop does (env.py:1318-1355: reset, then
`env.step(env.action_space.sample())`, reset again when an episode ends; no sleep, no render).

BUILD-CONTAINER ONLY (needs /root/reference); one process, one core.  Prints one JSON line per
case; the numbers quoted in BASELINE.md section 5 come from this script.
    python tools/ref_python_baseline.py [steps]
"""
import json
import os
import sys
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402


def run(num_humans, steps, seed):
    env = rh.make_env()
    env.env_param_range['num_humans'] = (list(num_humans), 'int')
    np.random.seed(seed)
    t_reset = time.perf_counter()
    env.reset()
    t_reset = time.perf_counter() - t_reset
    n_h, t_step, episodes = [len(env.humans)], 0.0, 1
    for _ in range(steps):
        a = env.action_space.sample()
        t0 = time.perf_counter()
        _, _, done, _ = env.step(a)
        t_step += time.perf_counter() - t0
        if done:
            t0 = time.perf_counter()
            env.reset()
            t_reset += time.perf_counter() - t0
            episodes += 1
            n_h.append(len(env.humans))
    rays = steps * 512 * (1 + float(np.mean(n_h)))
    return {"case": "num_humans in %s" % (list(num_humans),), "steps": steps, "episodes": episodes,
            "mean_humans": float(np.mean(n_h)),
            "env_steps_per_s": steps / t_step, "ms_per_step": 1e3 * t_step / steps,
            "robot_rays_per_s": steps * 512 / t_step, "all_rays_per_s": rays / t_step,
            "ms_per_reset": 1e3 * t_reset / episodes, "cores": 1,
            "what": "reference env.py unmodified (step only; resets timed separately), natives = C checker"}


if __name__ == '__main__':
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    # num_humans = 0 is not runnable: the reference feeds an empty batch to its policy and
    # human_policy.py:45 raises (view of 0 elements) -- 1 pedestrian is the smallest case
    for nh in ([1, 1], [5, 15]):
        print(json.dumps(run(nh, steps, seed=0)), flush=True)
