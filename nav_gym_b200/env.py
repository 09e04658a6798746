"""Single-environment drop-in for the reference's ``nav_gym_env.env`` module.

``gym.make('NavGym-v0')`` resolves to :class:`NavGymEnv` below: same constructor kwargs
(reference nav_gym_env/__init__.py:6-38, env.py:31-53), same ``reset() -> obs`` /
``step(action) -> (obs, reward, done, info)`` old-gym API with the reference's return layout
(env.py:455-462, 477-481, 728), same HER entry points (``compute_reward(s)``,
``compute_done`` / ``compute_terminals``, ``compute_info``, env.py:464-589) and the module
helpers ``ros_env.py`` imports.  Every per-step number comes from the CUDA kernels behind
include/navgym_b200.h (a batch of one environment); this module is host glue.

What differs from the reference, by design: episodes are sampled with this
package's map generator and spawn sampler (``maps.py``: maps in the reference's style, spawns
obeying env.py:379, 761, 779-783) rather than map_generator.py + pyastar2d; the pedestrians are
the reference's policy-driven ones (``pedestrians.PedestrianSim``, env.py:617-693) following
geodesic distance fields instead of pyastar2d paths, with random-init policy weights unless
NAVGYM_HUMAN_POLICY names the reference's undistributed ``human_policy.pth``.
"""
import numpy as np

from . import gym_shim, maps
from .robot import Human, KetiRobot

gym = gym_shim.install()
spaces = gym.spaces
EzPickle = gym.utils.EzPickle

DEFAULT_KWARGS = dict(
    robot_type='keti', time_step=0.2, min_turning_radius=0, distance_threshold=0.5,
    num_scan_stack=1, linvel_range=[0, 0.5], rotvel_range=[-0.64, 0.64],
    human_v_pref_range=[0., 0.6], human_has_legs_ratio=0.5, indoor_ratio=0.5,
    min_goal_dist=10, max_goal_dist=20, reward_scale=15., reward_success_factor=1,
    reward_crash_factor=1, reward_progress_factor=0.001, reward_forward_factor=0.0,
    reward_rotation_factor=0.005, reward_discomfort_factor=0.01,
    env_param_range=dict(
        num_humans=([5, 15], 'int'), corridor_width=([3, 4], 'int'), iterations=([80, 150], 'int'),
        obstacle_number=([10, 10], 'int'), obstacle_width=([0.3, 1.0], 'float'),
        scan_noise_std=([0., 0.05], 'float')),
)


# ---- module helpers (reference env.py:1214-1315) -------------------------------------------
def batch_ij_to_xy(ij, map_info):
    """Cell indices -> cell-centre coordinates (env.py:1214-1220)."""
    ij = np.asarray(ij)
    res, origin = map_info['resolution'], map_info['origin']
    return np.concatenate([(ij[..., 0:1] + 0.5) * res + origin[0],
                           (ij[..., 1:2] + 0.5) * res + origin[1]], axis=-1)


def ij_to_xy(ij, map_info):
    return batch_ij_to_xy(np.asarray(ij)[None, :], map_info)[0]


def batch_xy_to_ij(xy, map_info, clip_if_outside=True):
    """Coordinates -> int64 cell indices through a float32 intermediate, truncated, clipped
    against (height, width) in that order (env.py:1228-1253)."""
    xy = np.asarray(xy)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise IndexError("xy should be of shape (n, 2)")
    res, origin = map_info['resolution'], map_info['origin']
    ij = np.empty(xy.shape, np.float32)
    ij[:, 0] = (xy[:, 0].astype(np.float64) - origin[0]) / res
    ij[:, 1] = (xy[:, 1].astype(np.float64) - origin[1]) / res
    if clip_if_outside:
        ij[:, 0] = np.where(ij[:, 0] >= map_info['height'], map_info['height'] - 1, ij[:, 0])
        ij[:, 1] = np.where(ij[:, 1] >= map_info['width'], map_info['width'] - 1, ij[:, 1])
        ij = np.where(ij < 0, np.float32(0), ij)
    return ij.astype(np.int64)


def xy_to_ij(xy, map_info, clip_if_outside=True):
    return batch_xy_to_ij(np.asarray(xy)[None, :], map_info, clip_if_outside)[0]


def path_to_waypoints(path, interval):
    """Thin a path to points more than `interval` apart, ending on the last (env.py:1261-1277)."""
    path = np.asarray(path, np.float64)
    out, i = [], 0
    while True:
        far = np.where(np.linalg.norm(path[i:] - path[i], axis=-1) > interval)[0]
        if len(far) == 0:
            out.append(path[-1])
            break
        i += int(far[0])
        out.append(path[i])
    return np.array(out)


def observation_to_dict(observation, num_scan_stack, n_angles):
    n = num_scan_stack * n_angles
    other = observation[n:]
    return dict(scan_stack=observation[:n], scan=observation[n - n_angles:n], prev_pose=other[:2],
                pose=other[2:4], vel=other[4:6], yaw=other[6])


def observation_batch_to_dict(observation, num_scan_stack, n_angles):
    n = num_scan_stack * n_angles
    other = observation[:, n:]
    return dict(scan_stack=observation[:, :n], scan=observation[:, n - n_angles:n],
                prev_pose=other[:, :2], pose=other[:, 2:4], vel=other[:, 4:6], yaw=other[:, 6])


class NavGymEnv(gym.Env, EzPickle):
    def __init__(self, robot_type, time_step, min_turning_radius, distance_threshold, num_scan_stack,
                 linvel_range, rotvel_range, human_v_pref_range, human_has_legs_ratio, indoor_ratio,
                 min_goal_dist, max_goal_dist, reward_scale, reward_success_factor,
                 reward_crash_factor, reward_progress_factor, reward_forward_factor,
                 reward_rotation_factor, reward_discomfort_factor, env_param_range,
                 device='cuda:0'):
        EzPickle.__init__(self, robot_type, time_step, min_turning_radius, distance_threshold,
                          num_scan_stack, linvel_range, rotvel_range, human_v_pref_range,
                          human_has_legs_ratio, indoor_ratio, min_goal_dist, max_goal_dist,
                          reward_scale, reward_success_factor, reward_crash_factor,
                          reward_progress_factor, reward_forward_factor, reward_rotation_factor,
                          reward_discomfort_factor, env_param_range, device=device)
        if robot_type != 'keti':
            raise NotImplementedError
        self.robot_type, self.time_step = robot_type, time_step
        self.min_turning_radius, self.distance_threshold = min_turning_radius, distance_threshold
        self.num_scan_stack = num_scan_stack
        self.linvel_range, self.rotvel_range = linvel_range, rotvel_range
        self.human_v_pref_range, self.human_has_legs_ratio = human_v_pref_range, human_has_legs_ratio
        self.indoor_ratio = indoor_ratio
        self.min_goal_dist, self.max_goal_dist = min_goal_dist, max_goal_dist
        self.reward_scale = reward_scale
        self.reward_success_factor, self.reward_crash_factor = reward_success_factor, reward_crash_factor
        self.reward_progress_factor, self.reward_forward_factor = reward_progress_factor, reward_forward_factor
        self.reward_rotation_factor, self.reward_discomfort_factor = reward_rotation_factor, reward_discomfort_factor
        self.env_param_range = env_param_range
        self.device = device
        from .batched_env import scan_thresholds  # needs the CUDA library: fails loudly without it
        self.scan_threshold, self.scan_discomfort_threshold = scan_thresholds(device)
        self.prev_action = np.array([0., 0.])
        self.prev_obs = None
        self.env_param = None
        self.steps_since_reset = 0
        self.robot, self.humans, self.map_info, self._sim = None, [], None, None
        self.action_space = spaces.Box(low=np.array([linvel_range[0], rotvel_range[0]]),
                                       high=np.array([linvel_range[1], rotvel_range[1]]), dtype=np.float32)
        n = num_scan_stack * KetiRobot.n_angles + 7
        self.observation_space = spaces.Dict({
            'observation': spaces.Box(-np.inf, np.inf, shape=(n,), dtype=np.float32),
            'achieved_goal': spaces.Box(-np.inf, np.inf, shape=(2,), dtype=np.float32),
            'desired_goal': spaces.Box(-np.inf, np.inf, shape=(2,), dtype=np.float32)})

    def _reward_kwargs(self):
        return dict(reward_scale=self.reward_scale, reward_success_factor=self.reward_success_factor,
                    reward_crash_factor=self.reward_crash_factor,
                    reward_progress_factor=self.reward_progress_factor,
                    reward_forward_factor=self.reward_forward_factor,
                    reward_rotation_factor=self.reward_rotation_factor,
                    reward_discomfort_factor=self.reward_discomfort_factor)

    def _override_reward_factor(self, reward_scale=15., reward_success_factor=1, reward_crash_factor=1,
                                reward_progress_factor=0.001, reward_forward_factor=0.0,
                                reward_rotation_factor=0.005, reward_discomfort_factor=0.01):
        self.reward_scale = reward_scale
        self.reward_success_factor, self.reward_crash_factor = reward_success_factor, reward_crash_factor
        self.reward_progress_factor, self.reward_forward_factor = reward_progress_factor, reward_forward_factor
        self.reward_rotation_factor, self.reward_discomfort_factor = reward_rotation_factor, reward_discomfort_factor
        if self._sim is not None:
            a = self._sim.args
            a.r_scale, a.r_success, a.r_crash = reward_scale, reward_success_factor, reward_crash_factor
            a.r_progress, a.r_forward = reward_progress_factor, reward_forward_factor
            a.r_rotation, a.r_discomfort = reward_rotation_factor, reward_discomfort_factor

    # ---- episode sampling (host; reference env.py:281-383, 730-831) ---------------------
    def _sample_env_param(self):
        param = dict()
        for key, value in self.env_param_range.items():
            if value[1] == 'int':
                param[key] = np.random.choice(np.arange(value[0][0], value[0][1] + 1))
            elif value[1] == 'float':
                param[key] = np.random.uniform(value[0][0], value[0][1])
            else:
                raise NotImplementedError
        return param

    def reset(self):
        import torch
        from .batched_env import BatchedNavGym, filter_spawn_pool, MapPool
        self.env_param = self._sample_env_param()
        self.steps_since_reset = 0
        self.prev_action = np.array([0., 0.])
        self.prev_obs = None
        self.map_info, mp, pool = self._sample_world(MapPool, filter_spawn_pool)
        sx, sy, gx, gy, th = pool[np.random.randint(len(pool))]
        self.robot = KetiRobot(sx, sy, th, gx, gy, self.time_step)
        n_h = int(self.env_param['num_humans'])
        self._sim = BatchedNavGym(1, mp, device=self.device, time_step=self.time_step,
                                  distance_threshold=self.distance_threshold,
                                  min_turning_radius=self.min_turning_radius, early_stop=True,
                                  seed=int(np.random.randint(2 ** 31)),
                                  num_scan_stack=self.num_scan_stack, **self._reward_kwargs())
        self._sim.set_state([[sx, sy]], [[gx, gy]], [th],
                            noise_std=[self.env_param['scan_noise_std']])
        self._crowd = None
        if n_h:
            # the reference's pedestrians (env.py:617-693, 785-815): CNN policy on their own scans
            from .pedestrians import PedestrianSim
            self._crowd = PedestrianSim(self._sim, n_h, policy=self._human_policy(),
                                        v_pref_range=self.human_v_pref_range,
                                        has_legs_ratio=self.human_has_legs_ratio,
                                        seed=int(np.random.randint(2 ** 31)))
        self._sim.reset()
        torch.cuda.synchronize(self._sim.device)
        obs = self._obs()
        self.prev_obs = obs
        return obs

    def _sample_world(self, MapPool, filter_spawn_pool):
        """A map with its EDT on the device and a pool of valid (start, goal, heading) tuples
        (env.py:294-383, 748-783).  The reference draws a fresh map every episode, and so does
        this by default; NAVGYM_WORLD_CACHE=k keeps the last k worlds (map + EDT + spawn pool +
        the pedestrians' planning fields) and re-draws episodes among them once k exist, which
        takes the map generation, EDT build and BFS passes off all later resets."""
        import os
        keep = int(os.environ.get('NAVGYM_WORLD_CACHE', '0'))
        worlds = self.__dict__.setdefault('_worlds', [])
        if keep > 0 and len(worlds) >= keep:
            return worlds[np.random.randint(len(worlds))]
        for _ in range(20):
            if np.random.random() < self.indoor_ratio:
                map_info = maps.create_indoor_map(self.env_param['corridor_width'],
                                                  self.env_param['iterations'])
            else:
                map_info = maps.create_outdoor_map(self.env_param['obstacle_number'],
                                                   self.env_param['obstacle_width'])
            mp = MapPool([map_info], self.device)
            pool = maps.spawn_pool(map_info, 64, min_goal_dist=self.min_goal_dist,
                                   max_goal_dist=self.max_goal_dist)
            pool = filter_spawn_pool(map_info, pool, self.device, map_pool=mp)
            if len(pool):
                break
        else:
            raise RuntimeError('[sample_start_goal_path] something is wrong...')
        if keep > 0:
            worlds.append((map_info, mp, pool))
        return map_info, mp, pool

    _policy_cache = None

    def _human_policy(self):
        """The pedestrian policy network (env.py:112-118).  The reference loads
        nav_gym_env/human_policy.pth, a blob its repository does not distribute: set
        NAVGYM_HUMAN_POLICY to such a state_dict to use it, else the weights are random-init."""
        if NavGymEnv._policy_cache is None:
            import os
            import torch
            from .pedestrians import HumanPolicy
            pol = HumanPolicy()
            path = os.environ.get('NAVGYM_HUMAN_POLICY')
            if path:
                pol.load_state_dict(torch.load(path, map_location='cpu'))
            NavGymEnv._policy_cache = pol
        return NavGymEnv._policy_cache

    def _obs(self):
        """One packed device-to-host copy per step (BatchedNavGym.export_env)."""
        view = self._view = self._sim.export_env(0)
        self.robot.px, self.robot.py, self.robot.theta = view.robot.px, view.robot.py, view.robot.theta
        self.humans = view.humans
        return view.prev_obs

    def step(self, action):
        import torch
        self.steps_since_reset += 1
        action = np.array(action, dtype=np.float64)
        if action[0] < self.linvel_range[0] or action[0] > self.linvel_range[1]:
            print('linvel {} is out of range {}'.format(action[0], self.linvel_range))
        if action[1] < self.rotvel_range[0] or action[0] > self.rotvel_range[1]:  # sic, env.py:608
            print('rotvel {} is out of range {}'.format(action[1], self.rotvel_range))
        sim = self._sim
        if self._crowd is not None:
            self._crowd.step(torch.from_numpy(action[None].astype(np.float32)))
        else:
            sim.step(torch.from_numpy(action[None].astype(np.float32)))
        obs = self._obs()
        view = self._view
        self.robot.v, self.robot.r = float(action[0]), float(action[1])
        reward = np.float64(view.reward)
        done = np.bool_(view.done)
        info = {'is_success': np.float32(view.is_success), 'is_crash': np.float32(view.is_crash),
                'distance': np.float64(view.distance)}
        self.prev_action = action
        self.prev_obs = obs
        return obs, reward, done, info

    # ---- HER entry points (env.py:464-589), evaluated by the device kernel ----------------
    def _her(self, obs):
        import torch
        o = np.asarray(obs['observation'], np.float32)
        g = np.asarray(obs['desired_goal'], np.float32)
        single = o.ndim == 1
        if single:
            o, g = o[None], g[None]
        if self._sim is None:
            raise RuntimeError('reset() the environment first')
        out = self._sim.compute_rewards(torch.from_numpy(np.ascontiguousarray(o)),
                                        torch.from_numpy(np.ascontiguousarray(g)), **self._reward_kwargs())
        return {k: v.cpu().numpy() for k, v in out.items()}, single

    def compute_rewards(self, actions, obs, make_render_reward_txt=False):
        out, _ = self._her(obs)
        return out['reward'].astype(np.float64)

    def compute_reward(self, action, obs, make_render_reward_txt=False):
        return self.compute_rewards(None, {k: np.asarray(v)[None] for k, v in obs.items()})[0]

    def compute_terminals(self, obs):
        out, _ = self._her(obs)
        return out['done'].astype(bool)

    def compute_done(self, obs):
        return self.compute_terminals({k: np.asarray(v)[None] for k, v in obs.items()})[0]

    def compute_info(self, obs):
        out, _ = self._her({k: np.asarray(v)[None] for k, v in obs.items()})
        return {'is_success': np.float32(out['is_success'][0]), 'is_crash': np.float32(out['is_crash'][0]),
                'distance': np.float64(out['distance'][0])}

    def render(self, mode='human'):
        """Debug view (env.py:833-1058): 'human' opens the cv2 window, 'rgb_array' returns it."""
        from .render import render
        if self.map_info is None:
            raise RuntimeError('reset() the environment first')
        return render(self, mode)


def register_env():
    """Register 'NavGym-v0' with gym (the real package if importable, else the bundled shim),
    kwargs as in the reference's nav_gym_env/__init__.py:4-40."""
    reg = getattr(gym.envs, 'registration', None)
    registry = getattr(reg, 'registry', None)
    try:
        known = 'NavGym-v0' in registry or 'NavGym-v0' in getattr(registry, 'env_specs', {})
    except TypeError:
        known = False
    if not known:
        gym.envs.registration.register(id='NavGym-v0', kwargs=dict(DEFAULT_KWARGS),
                                       entry_point='nav_gym_b200.env:NavGymEnv')
