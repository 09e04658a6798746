"""Policy-driven pedestrians on the device (reference env.py:617-693).

The reference moves every pedestrian with a small CNN policy (human_policy.py:19-71) fed with
the pedestrian's own 512-beam 180 deg / 6 m lidar scan, its local goal and its previous action,
then re-scans the world from each pedestrian's new pose.  Here the same pipeline runs batched
over [num_envs, max_ped] on the GPU:

  agent_scans()            navgym_agent_scan_batch: every pedestrian's scan in one launch
  HumanPolicy              the reference architecture in torch (the one dense-math component;
                           its pretrained weights, human_policy.pth, are not distributed: load
                           them with load_state_dict when available, else random-init)
  PedestrianSim            device state + the per-step sequence of env.py:617-693

PyTorch carries the dense policy forward and the elementwise bookkeeping; raycasts are the
library's CUDA kernels.  Nothing here touches the CPU oracle.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .robot import Human, KetiRobot, beam_table


class HumanPolicy(nn.Module):
    """Actor / critic of human_policy.py:19-71 with the reference's parameter names and
    construction order, so that its state_dict loads here and `torch.manual_seed(s);
    HumanPolicy()` draws the same initial weights as the reference class does."""

    def __init__(self, frames=3, action_space=2):
        super().__init__()
        self.logstd = nn.Parameter(torch.zeros(action_space))
        self.act_fea_cv1 = nn.Conv1d(frames, 32, kernel_size=5, stride=2, padding=1)
        self.act_fea_cv2 = nn.Conv1d(32, 32, kernel_size=3, stride=2, padding=1)
        self.act_fc1 = nn.Linear(128 * 32, 256)
        self.act_fc2 = nn.Linear(256 + 2 + 2, 128)
        self.actor1 = nn.Linear(128, 1)
        self.actor2 = nn.Linear(128, 1)
        self.crt_fea_cv1 = nn.Conv1d(frames, 32, kernel_size=5, stride=2, padding=1)
        self.crt_fea_cv2 = nn.Conv1d(32, 32, kernel_size=3, stride=2, padding=1)
        self.crt_fc1 = nn.Linear(128 * 32, 256)
        self.crt_fc2 = nn.Linear(256 + 2 + 2, 128)
        self.critic = nn.Linear(128, 1)

    def mean(self, x, goal, speed):
        """The deterministic action mean, the only output env.py:652-656 uses.
        x [N, frames, 512] preprocessed scans, goal [N, 2] local goal, speed [N, 2]."""
        a = F.relu(self.act_fea_cv1(x))
        a = F.relu(self.act_fea_cv2(a))
        a = F.relu(self.act_fc1(a.reshape(a.shape[0], -1)))
        a = F.relu(self.act_fc2(torch.cat((a, goal, speed), dim=-1)))
        return torch.cat((torch.sigmoid(self.actor1(a)), torch.tanh(self.actor2(a))), dim=-1)

    def forward(self, x, goal, speed):
        """(value, sampled action, log-prob, mean) as the reference's forward."""
        mean = self.mean(x, goal, speed)
        logstd = self.logstd.expand_as(mean)
        std = torch.exp(logstd)
        action = torch.normal(mean, std)
        logprob = (-(action - mean).pow(2) / (2 * std.pow(2)) - 0.5 * np.log(2 * np.pi) - logstd).sum(-1, keepdim=True)
        v = F.relu(self.crt_fea_cv1(x))
        v = F.relu(self.crt_fea_cv2(v))
        v = F.relu(self.crt_fc1(v.reshape(v.shape[0], -1)))
        v = self.critic(F.relu(self.crt_fc2(torch.cat((v, goal, speed), dim=-1))))
        return v, action, logprob, mean


def preprocess_scan(scan):
    """env.py:627-629, 648: clip to the pedestrian lidar's 6 m and centre -- in float64, as the
    reference does before its `.float()`."""
    return (torch.clamp(scan.double(), 0.0, Human.range_max) / Human.range_max - 0.5).float()


def footprint_polygons(pose, footprint):
    """Closed world-frame footprints as segments (env.py:404-414): pose [..., 3] float64 ->
    float32 [..., 4, 4] rows (ax, ay, bx, by)."""
    fp = torch.as_tensor(footprint, dtype=torch.float64, device=pose.device)  # [4, 2]
    c, s = torch.cos(pose[..., 2:3]), torch.sin(pose[..., 2:3])
    wx = c * fp[:, 0] - s * fp[:, 1] + pose[..., 0:1]
    wy = s * fp[:, 0] + c * fp[:, 1] + pose[..., 1:2]
    pts = torch.stack((wx, wy), dim=-1).to(torch.float32)                     # [..., 4, 2]
    return torch.cat((pts, torch.roll(pts, -1, dims=-2)), dim=-1)


def human_set_vel(pose, action, dt):
    """Human.set_vel (human.py:32-41) on float64 tensors: pose [..., 3], action [..., 2] =
    (linvel, rotvel) -> (new pose, world velocity (vx, vy) = linvel * (cos, sin)(old theta))."""
    x, y, th = pose[..., 0], pose[..., 1], pose[..., 2]
    v, w = action[..., 0], action[..., 1]
    vel = torch.stack((v * torch.cos(th), v * torch.sin(th)), dim=-1)
    th1 = th + w * dt
    x = x + torch.cos(th1) * v * dt
    y = y + torch.sin(th1) * v * dt
    return torch.stack((x, y, torch.remainder(th1, 2 * np.pi)), dim=-1), vel


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class AgentScanner(object):
    """navgym_agent_scan_batch bound to one MapPool: scans of [num_envs, agents_per_env] agents
    with the pedestrian lidar (human.py:11-16) against the map and the other agents' closed
    footprints."""

    def __init__(self, pool, num_envs, agents_per_env, max_seg, map_id, agent=Human, cell_rule='numpy1'):
        self.lib = _lib.load()
        self.pool, self.device = pool, pool.device
        self.B, self.P, self.K = int(num_envs), int(agents_per_env), int(agent.n_angles)
        self.lin = torch.from_numpy(beam_table(agent)).to(self.device)
        self.ranges = torch.zeros(self.B, self.P, self.K, dtype=torch.float32, device=self.device)
        a = _lib.ScanArgs()
        a.num_envs, a.agents_per_env, a.num_beams, a.max_seg = self.B, self.P, self.K, int(max_seg)
        a.cell_rule = {'numpy1': 0, 'numpy2': 1}[cell_rule]
        a.range_max = float(agent.range_max)
        res = float(pool.maps[0]['resolution'])
        a.t_stop = float(agent.range_max) / res + 2.0  # farther samples are clipped anyway
        a.maps, a.edt_pool = _ptr(pool.maps_dev), _ptr(pool.edt_pool)
        self.map_id = map_id
        a.map_id, a.lin, a.ranges = _ptr(map_id), _ptr(self.lin), _ptr(self.ranges)
        self.args = a

    def scan(self, pose, segs, nseg, skip=None, nagent=None):
        """pose f64 [B, P, 3]; segs f32 [B, max_seg, 4]; nseg i32 [B]; skip i32 [B, P, 2] (each
        agent's own segments) -> ranges f32 [B, P, K] (owned by the scanner, overwritten)."""
        a = self.args
        assert pose.dtype == torch.float64 and pose.is_contiguous() and tuple(pose.shape) == (self.B, self.P, 3)
        assert segs.dtype == torch.float32 and segs.is_contiguous() and segs.shape[1] == a.max_seg
        self._keep = (pose, segs, nseg, skip, nagent)
        a.pose, a.segs, a.nseg, a.skip, a.nagent = _ptr(pose), _ptr(segs), _ptr(nseg), _ptr(skip), _ptr(nagent)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.navgym_agent_scan_batch(C.byref(a), C.c_void_p(stream)), 'agent_scan_batch')
        return self.ranges


class PedestrianSim(object):
    """The reference's pedestrians for a whole BatchedNavGym, on the device.

    Per step, in the reference's order (env.py:617-693):
        sim.act()            route following -> policy -> Human.set_vel -> leg odometry ->
                             geometry the robot's lidar sees (discs for legs, boxes otherwise)
        env.step(actions)    the robot's fused step, scanning against that geometry
        sim.observe()        every pedestrian's own scan from its new pose (robot included),
                             the next policy input
    `sim.step(actions)` does the three.  Environments whose episode ended respawn their
    pedestrians (env.py:785-806) before the next act().

    State, all [num_envs, max_ped, ...] device tensors: pose f64 (x, y, theta), vel f64 (vx, vy),
    v_pref f64, has_legs bool, dist_travelled f64 (x, y, theta in the base frame), prev_yaw f64,
    prev_action f32 (the clipped policy mean, the policy's `speed` input), scan f32 [.., 512],
    goal_id i32, waypoint f64.
    """

    def __init__(self, env, max_ped, nped=None, policy=None, v_pref_range=(0.0, 0.6), has_legs_ratio=0.5,
                 num_goals=32, min_goal_dist=10.0, min_robot_dist=4.0, seed=0, fold_frames=True):
        from . import maps as M
        self.env, self.device = env, env.device
        self.B, self.P = env.B, int(max_ped)
        self.dt = float(env.args.dt)
        self.v_pref_range, self.has_legs_ratio = v_pref_range, float(has_legs_ratio)
        self.min_robot_dist = float(min_robot_dist)
        self.lib = _lib.load()
        dev, f64, f32, i32 = self.device, torch.float64, torch.float32, torch.int32
        B, P = self.B, self.P
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(int(seed))
        self.nped = None if nped is None else torch.as_tensor(nped, dtype=i32).to(dev)
        self.live = (torch.arange(P, device=dev)[None, :] < (self.nped[:, None] if self.nped is not None else P))
        if policy is None:
            policy = HumanPolicy()
        self.policy = policy.to(dev).eval()
        self.fold_frames = bool(fold_frames)
        # ---- planning fields per map
        rng = np.random.RandomState(int(seed) + 77)
        pm = (_lib.PlanMapT * len(env.pool.maps))()
        fl, gl, self.free_xy, self.free_off = [], [], [], [0]
        foff = goff = 0
        for i, m in enumerate(env.pool.maps):
            fields, goals, cm = M.goal_fields(m, num_goals, rng)
            pm[i] = _lib.PlanMapT(cm['width'], cm['height'], num_goals, 0, foff, goff,
                                  float(cm['origin'][0]), float(cm['origin'][1]), float(cm['resolution']))
            fl.append(fields.reshape(-1))
            gl.append(goals)
            foff += fields.size
            goff += num_goals
            r, c = np.where(cm['data'] == 0)
            xy = np.column_stack([(c + 0.5) * cm['resolution'] + cm['origin'][0],
                                  (r + 0.5) * cm['resolution'] + cm['origin'][1]])
            self.free_xy.append(xy)
            self.free_off.append(self.free_off[-1] + len(xy))
        self.plan_maps = torch.from_numpy(np.frombuffer(bytes(pm), dtype=np.uint8).copy()).to(dev)
        self.fields = torch.from_numpy(np.concatenate(fl).view(np.int16)).to(dev)
        self.goals = torch.from_numpy(np.concatenate(gl)).to(dev)
        self.free_xy = torch.from_numpy(np.concatenate(self.free_xy)).to(dev)
        self.free_off = torch.tensor(self.free_off, dtype=torch.int64, device=dev)
        self.num_goals = int(num_goals)
        # ---- state
        self.pose = torch.zeros(B, P, 3, dtype=f64, device=dev)
        self.vel = torch.zeros(B, P, 2, dtype=f64, device=dev)
        self.v_pref = torch.zeros(B, P, dtype=f64, device=dev)
        self.has_legs = torch.zeros(B, P, dtype=torch.bool, device=dev)
        self.dist_travelled = torch.zeros(B, P, 3, dtype=f64, device=dev)
        self.prev_yaw = torch.zeros(B, P, dtype=f64, device=dev)
        self.prev_action = torch.zeros(B, P, 2, dtype=f32, device=dev)
        self.goal_id = torch.zeros(B, P, dtype=i32, device=dev)
        self.waypoint = torch.full((B, P, 2), float('nan'), dtype=f64, device=dev)
        self.goal_local = torch.zeros(B, P, 2, dtype=f32, device=dev)
        a = _lib.PlanArgs()
        a.num_envs, a.max_ped, a.step, a.seed = B, P, 0, int(seed)
        a.env_offset, a.min_goal_dist = int(env.args.env_offset), float(min_goal_dist)
        a.maps, a.fields, a.goals = _ptr(self.plan_maps), _ptr(self.fields), _ptr(self.goals)
        a.map_id, a.nped, a.pose = _ptr(env.map_id), _ptr(self.nped), _ptr(self.pose)
        a.goal_id, a.waypoint, a.goal_local = _ptr(self.goal_id), _ptr(self.waypoint), _ptr(self.goal_local)
        self.plan_args = a
        # ---- lidar of the pedestrians: segments [robot | ped 0 | ped 1 ...], 4 each
        self.scanner = AgentScanner(env.pool, B, P, 4 * (P + 1), env.map_id,
                                    cell_rule='numpy2' if env.args.cell_rule else 'numpy1')
        self.segs = torch.zeros(B, 4 * (P + 1), 4, dtype=f32, device=dev)
        nl = self.nped if self.nped is not None else torch.full((B,), P, dtype=i32, device=dev)
        self.nseg = (4 * (nl + 1)).to(i32)
        sk = torch.tensor([[4 * (1 + j), 4] for j in range(P)], dtype=i32, device=dev)
        self.skip = sk[None].expand(B, P, 2).contiguous()
        self.scan = self.scanner.ranges
        # ---- geometry for the robot's lidar goes through the env's pedestrian rows
        rows = torch.zeros(B, P, _lib.PED_F, dtype=f32, device=dev)
        env.attach_pedestrians(rows, self.nped)
        env.peds_scripted = False  # the sim moves them; env.step only emits their geometry
        self.reset()

    # ---- episode start (env.py:785-815) ------------------------------------------------
    def reset(self, mask=None):
        """(Re)spawn the pedestrians of the masked environments (all when None): a free cost-map
        cell at least 4 m from the robot (env.py:369-373), random heading, preferred speed and
        legs (env.py:793-803), a goal field, zero odometry; then their first scans."""
        B, P, dev = self.B, self.P, self.device
        if mask is None:
            mask = torch.ones(B, dtype=torch.bool, device=dev)
        m2 = mask[:, None].expand(B, P)
        mid = self.env.map_id.long()
        lo, n = self.free_off[mid], (self.free_off[mid + 1] - self.free_off[mid])
        rob = self.env.state[:2].t()  # [B, 2]
        xy = torch.zeros(B, P, 2, dtype=torch.float64, device=dev)
        todo = torch.ones(B, P, dtype=torch.bool, device=dev)
        for _ in range(6):  # rejection sampling, vectorised; the last draw is kept regardless
            u = torch.rand(B, P, generator=self.gen, device=dev, dtype=torch.float64)
            idx = lo[:, None] + torch.clamp((u * n[:, None]).long(), max=(n[:, None] - 1).clamp(min=0))
            cand = self.free_xy[idx]
            xy = torch.where(todo[..., None], cand, xy)
            todo = todo & ((cand - rob[:, None, :]).norm(dim=-1) < self.min_robot_dist)
        th = torch.rand(B, P, generator=self.gen, device=dev, dtype=torch.float64) * (2 * np.pi)
        vp = self.v_pref_range[0] + torch.rand(B, P, generator=self.gen, device=dev, dtype=torch.float64) * (
            self.v_pref_range[1] - self.v_pref_range[0])
        legs = torch.rand(B, P, generator=self.gen, device=dev) < self.has_legs_ratio
        gid = torch.randint(self.num_goals, (B, P), generator=self.gen, device=dev, dtype=torch.int32)
        self.pose = torch.where(m2[..., None], torch.cat((xy, th[..., None]), -1), self.pose).contiguous()
        self.plan_args.pose = _ptr(self.pose)
        self.v_pref = torch.where(m2, vp, self.v_pref)
        self.has_legs = torch.where(m2, legs, self.has_legs)
        self.goal_id.copy_(torch.where(m2, gid, self.goal_id))
        self.waypoint.copy_(torch.where(m2[..., None], torch.full_like(self.waypoint, float('nan')), self.waypoint))
        z3 = torch.zeros_like(self.dist_travelled)
        self.dist_travelled = torch.where(m2[..., None], z3, self.dist_travelled)
        self.vel = torch.where(m2[..., None], torch.zeros_like(self.vel), self.vel)
        self.prev_action = torch.where(m2[..., None], torch.zeros_like(self.prev_action), self.prev_action)
        self._emit()
        self.observe()

    # ---- env.py:617-662 + 664-681 + 237-255 ---------------------------------------------
    @torch.no_grad()
    def act(self):
        env, a = self.env, self.plan_args
        if env.args.auto_reset:  # episodes that ended in the last step start with new pedestrians
            done = env.done.bool()
            self.reset(done)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.navgym_peds_plan(C.byref(a), env._stream()), 'peds_plan')
        a.step += 1
        x = preprocess_scan(self.scan).reshape(self.B * self.P, 1, -1)
        goal, speed = self.goal_local.reshape(-1, 2), self.prev_action.reshape(-1, 2)
        if self.fold_frames:
            mean = self._mean_folded(x, goal, speed)
        else:
            mean = self.policy.mean(x.expand(-1, 3, -1).contiguous(), goal, speed)
        lo = torch.tensor([0.0, -1.0], device=self.device)
        hi = torch.tensor([1.0, 1.0], device=self.device)
        act = torch.minimum(torch.maximum(mean, lo), hi).reshape(self.B, self.P, 2)   # env.py:655-657
        self.prev_action = act
        action = act.double() * self.v_pref[..., None]                                # env.py:659-661
        self.prev_yaw = torch.atan2(torch.sin(self.pose[..., 2]), torch.cos(self.pose[..., 2]))
        pose, vel = human_set_vel(self.pose, action, self.dt)
        live = self.live[..., None]
        self.pose.copy_(torch.where(live, pose, self.pose))
        self.vel = torch.where(live, vel, self.vel)
        self._update_dist_travelled()
        self._emit()

    def _mean_folded(self, x, goal, speed):
        """env.py:647 hands the newest scan to all three input frames, so the first convolution
        sees three identical channels: sum its weights over them once and convolve one."""
        p = self.policy
        w = p.act_fea_cv1.weight.sum(dim=1, keepdim=True)
        h = F.relu(F.conv1d(x, w, p.act_fea_cv1.bias, stride=2, padding=1))
        h = F.relu(p.act_fea_cv2(h))
        h = F.relu(p.act_fc1(h.reshape(h.shape[0], -1)))
        h = F.relu(p.act_fc2(torch.cat((h, goal, speed), dim=-1)))
        return torch.cat((torch.sigmoid(p.actor1(h)), torch.tanh(p.actor2(h))), dim=-1)

    def _update_dist_travelled(self):
        """env.py:237-255 with pose2d's inverse_pose2d / apply_tf_to_vel written out: the world
        velocity rotated into the base frame, integrated."""
        th = self.pose[..., 2]
        vrot = (th - self.prev_yaw) / self.dt
        c, s = torch.cos(th), torch.sin(th)
        vx, vy = self.vel[..., 0], self.vel[..., 1]
        base = torch.stack((c * vx + s * vy, -s * vx + c * vy, vrot), dim=-1)
        self.dist_travelled = self.dist_travelled + torch.where(self.live[..., None], base * self.dt, torch.zeros_like(base))

    def _emit(self):
        """Pedestrian rows of the env (layout include/navgym_b200.h) -> discs / segments."""
        rows = self.env.peds
        rows[..., 0:3] = self.pose.float()
        rows[..., 9:12] = self.dist_travelled.float()
        rows[..., 12] = self.has_legs.float()
        self.env._peds_emit(advance=False)

    # ---- env.py:683-693 -------------------------------------------------------------------
    @torch.no_grad()
    def observe(self):
        """Every pedestrian's scan from its current pose: the map, the robot's threshold
        footprint and the other pedestrians' footprints (lidar_legs=False), no noise."""
        rob = self.env.state[:3].t().contiguous()  # [B, 3] px, py, theta
        self.segs[:, :4] = footprint_polygons(rob, KetiRobot.threshold_footprint)
        self.segs[:, 4:] = footprint_polygons(self.pose, Human.footprint).reshape(self.B, -1, 4)
        self.scanner.scan(self.pose, self.segs, self.nseg, self.skip, self.nped)
        return self.scan

    def step(self, actions):
        self.act()
        out = self.env.step(actions)
        self.observe()
        return out
