"""GPU: the fused step at BASELINE.json's full size (C2: 4096 environments on the bench world)
through properties that do not need the oracle to finish the whole batch -- independence of the
launch order and of the sharding, invariants of the outputs, the oracle on a random subsample --
and the edge cases of the boundary (empty batch, robot inside an obstacle or on the map border,
full and empty obstacle lists)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as orc

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
B_FULL = 4096


@pytest.fixture(scope='module')
def world():
    from bench import build_world
    from nav_gym_b200.batched_env import MapPool, filter_spawn_pool
    m, pool = build_world(0, 16384)
    pool = filter_spawn_pool(m, pool, 'cuda:0')
    return m, pool, MapPool([m], 'cuda:0', spawn_pools=[pool])


def _env(world, B, rows, sigma, **kw):
    from nav_gym_b200.batched_env import BatchedNavGym
    m, pool, mp = world
    env = BatchedNavGym(B, mp, seed=11, auto_reset=True, **kw)
    env.set_state(rows[:, 0:2], rows[:, 2:4], rows[:, 4], noise_std=sigma)
    env.reset()
    return env


def _actions(T, B, seed):
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    return torch.rand(T, B, 2, device='cuda', generator=g) * torch.tensor([0.5, 1.28], device='cuda') + torch.tensor([0, -0.64], device='cuda')


def test_full_size_is_independent_of_launch_order_and_sharding(world):
    """4096 environments, 40 steps with auto-reset and Philox noise: (a) the longest-first CTA
    order changes nothing, (b) two shards of 2048 with env_offset reproduce the single batch
    bit for bit -- noise and respawns are keyed by the global environment id."""
    m, pool, mp = world
    rng = np.random.RandomState(3)
    rows = pool[rng.randint(len(pool), size=B_FULL)]
    sigma = rng.uniform(0, 0.05, B_FULL).astype(np.float32)
    T = 40
    acts = _actions(T, B_FULL, 1)
    ref = _env(world, B_FULL, rows, sigma, longest_first=True)
    plain = _env(world, B_FULL, rows, sigma, longest_first=False)
    half = B_FULL // 2
    sh = [_env(world, half, rows[i * half:(i + 1) * half], sigma[i * half:(i + 1) * half], env_offset=i * half)
          for i in range(2)]
    n_done = 0
    for t in range(T):
        ref.step(acts[t])
        plain.step(acts[t])
        for i in range(2):
            sh[i].step(acts[t, i * half:(i + 1) * half])
        torch.cuda.synchronize()
        for name in ('obs', 'reward', 'done', 'is_success', 'is_crash', 'distance', 'steps', 'state'):
            a = getattr(ref, name)
            assert torch.equal(a, getattr(plain, name)), (t, name, 'launch order')
            cat = torch.cat([getattr(s, name) for s in sh], dim=1 if name == 'state' else 0)
            assert torch.equal(a, cat), (t, name, 'sharding')
        n_done += int(ref.done.sum())
    assert n_done > 100  # episodes did end and respawn along the way


def test_full_size_invariants_and_oracle_subsample(world):
    """4096 environments without noise or auto-reset: output invariants on every row, and the
    oracle's lockstep batch on 96 randomly chosen environments, bit-exact, for 12 steps."""
    from nav_gym_b200.batched_env import BatchedNavGym
    m, pool, mp = world
    rng = np.random.RandomState(4)
    rows = pool[rng.randint(len(pool), size=B_FULL)]
    env = BatchedNavGym(B_FULL, mp, seed=1, auto_reset=False, record_hits=True)
    env.set_state(rows[:, 0:2], rows[:, 2:4], rows[:, 4], noise_std=np.zeros(B_FULL, np.float32))
    env.reset()
    pick = np.sort(rng.choice(B_FULL, 96, replace=False))
    o = orc.OracleBatch([m], np.zeros(96, np.int32), rows[pick, 0:2], rows[pick, 2:4], rows[pick, 4],
                        params=dict(t_stop=502.0))
    o.reset_obs()
    torch.cuda.synchronize()
    assert np.array_equal(env.obs[pick, :512].cpu().numpy(), o.obs[:, :512])
    acts = _actions(12, B_FULL, 2)
    thr = torch.from_numpy(env.scan_threshold).cuda()
    for t in range(12):
        env.step(acts[t])
        o.step(acts[t, pick].cpu().numpy())
        torch.cuda.synchronize()
        scan = env.obs[:, :512]
        assert float(scan.min()) >= 0.0 and float(scan.max()) <= 25.0
        assert torch.isfinite(env.obs).all() and torch.isfinite(env.reward).all()
        crash, succ, done = env.is_crash.bool(), env.is_success.bool(), env.done.bool()
        assert torch.equal(done, crash | succ)                      # env.py:507: no time limit here
        assert torch.equal(succ, env.distance < 0.5)
        # a crashed environment was rolled back: pose == previous pose in the returned row
        rb = env.obs[crash]
        assert torch.equal(rb[:, 512:514], rb[:, 514:516])
        # hit cells and ranges agree: range = |hit - origin| * 0.05 where there is a hit
        hits = env.hits.view(B_FULL, 512, 2).float()
        none = env.hits.view(B_FULL, 512, 2)[..., 0] == -32768
        r = torch.sqrt(hits[..., 0] ** 2 + hits[..., 1] ** 2) * 0.05
        ok = ~none & ~crash[:, None] & (r < 25.0)                   # (crashed rows hold the re-scan)
        assert torch.allclose(scan[ok], r[ok], atol=1e-5)
        assert np.array_equal(env.obs[pick, :512].cpu().numpy(), o.obs[:, :512]), t
        assert np.array_equal(env.done[pick].cpu().numpy(), o.done), t
        assert np.array_equal(env.is_crash[pick].cpu().numpy(), o.is_crash), t
        assert np.allclose(env.reward[pick].cpu().numpy(), o.reward, rtol=1e-6, atol=2e-6), t
    assert int(env.is_crash.sum()) >= 0 and thr.numel() == 512


def test_edge_cases_match_oracle():
    """Robot inside an obstacle (origin cell occupied: every beam ends at range 0), robot on the
    map border looking out, robot outside the map (origin clipped into it, env.py:1246-1253);
    obstacle lists empty and full (64 discs, 128 segments)."""
    import cuda_util
    import synth
    rng = np.random.RandomState(9)
    m = synth.outdoor_map(rng, size=400, n_obs=10)
    occ = np.asarray(m['data']) >= 0.1
    ys, xs = np.where(occ[6:-6, 6:-6])
    inside = np.array([(xs[0] + 6 + 0.5) * 0.05, (ys[0] + 6 + 0.5) * 0.05])
    start = np.array([inside, [0.26, 10.0], [19.74, 10.0], [10.0, 0.26], [-0.5, 5.0], [21.0, 25.0], [10.0, 10.0], [5.0, 5.0]])
    B = len(start)
    goal = start + 3.0
    theta = np.array([0.3, np.pi, 0.0, -np.pi / 2, 1.0, 2.0, 0.0, 4.0])
    md, ms = 64, 128
    map_id = np.zeros(B, np.int32)
    o = orc.OracleBatch([m], map_id, start, goal, theta, params=dict(t_stop=502.0), max_disc=md, max_seg=ms)
    c = cuda_util.CudaStepper([m], map_id, start, goal, theta, max_disc=md, max_seg=ms, early_stop=True)
    discs = np.zeros((B, md, 3), np.float32)
    segs = np.zeros((B, ms, 4), np.float32)
    discs[..., :2] = start[:, None, :] + rng.uniform(-5, 5, (B, md, 2))
    discs[..., 2] = 0.2
    a = start[:, None, :] + rng.uniform(-5, 5, (B, ms, 2))
    segs[..., :2], segs[..., 2:] = a, a + rng.uniform(-0.5, 0.5, (B, ms, 2))
    ndisc = np.array([0, md, 0, md, 1, md, 0, 17], np.int32)
    nseg = np.array([0, 0, ms, ms, 1, ms, 0, 33], np.int32)
    for stepper in (o, c):
        stepper.reset_obs(discs, ndisc, segs, nseg)
    assert np.array_equal(c.obs[:, :512], o.obs[:, :512])
    assert (c.obs[0, :512] == 0).all()                                # inside an obstacle
    for t in range(4):
        act = rng.uniform([0, -0.6], [0.5, 0.6], (B, 2)).astype(np.float32)
        for stepper in (o, c):
            stepper.step(act, discs, ndisc, segs, nseg)
        assert np.array_equal(c.obs[:, :512], o.obs[:, :512]), t
        assert np.array_equal(c.hits, o.hits), t
        assert np.array_equal(c.done, o.done) and np.array_equal(c.is_crash, o.is_crash), t
        assert np.allclose(c.reward, o.reward, rtol=1e-6, atol=2e-6), t


def test_empty_batch_is_a_no_op():
    from nav_gym_b200 import _lib
    lib = _lib.load()
    a = _lib.StepArgs()
    a.num_envs = 0
    assert lib.navgym_step_batch(C.byref(a), None) == 0
    assert lib.navgym_reset_obs_batch(C.byref(a), None) == 0
    h = _lib.HerArgs()
    assert lib.navgym_compute_rewards(C.byref(h), None) == 0
    p = _lib.PedsArgs()
    assert lib.navgym_peds_advance(C.byref(p), None) == 0
    for args, fn in ((_lib.ScanArgs(), lib.navgym_agent_scan_batch), (_lib.PlanArgs(), lib.navgym_peds_plan),
                     (_lib.MoveArgs(), lib.navgym_peds_move)):
        assert fn(C.byref(args), None) == 0
    # malformed: an observation stride too small for the row is refused, not launched
    a.num_envs, a.obs_stride = 4, 100
    assert lib.navgym_step_batch(C.byref(a), None) != 0
