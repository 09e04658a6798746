"""Latency of the single-environment drop-in (BASELINE config 1: gym.make('NavGym-v0'), random
actions, the reference's default kwargs): seconds per reset() and env-steps/s of step(), with the
reference's 5-15 policy-driven pedestrians and without pedestrians.  The reference's own Python
loop on one core of the build container (tools/ref_python_baseline.py):
58 env-steps/s with 5-15 pedestrians, 372 with one."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nav_gym_b200  # noqa: F401  (registers NavGym-v0; installs the gym shim when gym is absent)
import gym
from nav_gym_b200.env import DEFAULT_KWARGS


def run(label, kwargs, episodes=6, steps=300):
    np.random.seed(0)
    env = gym.make('NavGym-v0', **kwargs)
    t_reset, t_step, n_step = [], 0.0, 0
    for ep in range(episodes):
        t0 = time.perf_counter()
        env.reset()
        t_reset.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(steps):
            obs, r, d, info = env.step(env.action_space.sample())
            n_step += 1
            if d:
                break
        t_step += time.perf_counter() - t0
    return {"case": label, "episodes": episodes, "steps": n_step,
            "reset_s_first": t_reset[0], "reset_s_median": float(np.median(t_reset[1:])),
            "env_steps_per_s": n_step / t_step, "ms_per_step": 1e3 * t_step / n_step}


if __name__ == '__main__':
    rng = dict(DEFAULT_KWARGS['env_param_range'])
    none = dict(rng); none['num_humans'] = ([0, 0], 'int')
    print(json.dumps(run('default (5-15 policy-driven pedestrians)', {})))
    print(json.dumps(run('no pedestrians', {'env_param_range': none})))
    os.environ['NAVGYM_WORLD_CACHE'] = '4'
    print(json.dumps(run('default, NAVGYM_WORLD_CACHE=4', {}, episodes=10)))
