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
