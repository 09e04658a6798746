"""GPU: the drop-in surface (gym.make('NavGym-v0')), the HER batch kernel against the
reference's recorded rewards, and the scripted-pedestrian kernel."""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', gu.trace_names())
def test_her_kernel_reproduces_reference_rewards(name):
    """Stored observations of a reference trace -> compute_rewards / terminals / info on the
    device == what the reference's own compute_reward / compute_done / compute_info returned."""
    from nav_gym_b200.batched_env import BatchedNavGym
    G = gu.load(name)
    S = int(G['num_scan_stack']) if 'num_scan_stack' in G else 1
    env = BatchedNavGym(1, [gu.map_info(G)], device='cuda:0', num_scan_stack=S, **gu.env_kwargs(G))
    assert np.array_equal(env.scan_threshold, G['thr'])
    obs = np.concatenate([G['scan_stack'] if S > 1 else G['scan'], G['tail'].astype(np.float32)], axis=1)
    # on a crash the reference returns the re-scanned observation, whose reward terms were
    # computed on the crashed one: compare the steps whose returned obs is the scored obs
    keep = G['is_crash'] == 0
    out = env.compute_rewards(torch.from_numpy(obs).cuda(), torch.from_numpy(G['desired'].astype(np.float32)).cuda())
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert np.array_equal(out['is_success'][keep], G['is_success'][keep])
    assert np.array_equal(out['is_crash'][keep], G['is_crash'][keep])
    assert np.array_equal(out['done'][keep], G['done'][keep])
    assert np.allclose(out['reward'][keep], G['reward'][keep], rtol=0, atol=2e-5)
    assert np.allclose(out['distance'][keep], G['distance'][keep], rtol=0, atol=1e-5)


def test_gym_make_dropin_contract():
    import nav_gym_b200  # noqa: F401  (registers NavGym-v0)
    from nav_gym_b200 import gym_shim
    gym = gym_shim.install()
    np.random.seed(3)
    env = gym.make('NavGym-v0')
    assert env.action_space.shape == (2,) and env.observation_space['observation'].shape == (519,)
    obs = env.reset()
    assert set(obs) == {'observation', 'achieved_goal', 'desired_goal'}
    assert obs['observation'].shape == (519,) and obs['observation'].dtype == np.float64
    assert np.array_equal(obs['observation'][512:514], obs['observation'][514:516])  # prev_pose == pose
    assert np.array_equal(obs['observation'][516:518], [0, 0])
    d0 = np.linalg.norm(obs['achieved_goal'] - obs['desired_goal'])
    assert env.min_goal_dist < d0 < env.max_goal_dist
    assert 5 <= len(env.humans) <= 15
    prev = obs
    for t in range(25):
        a = np.array([0.3, 0.2 * np.sin(t)])
        obs, reward, done, info = env.step(a)
        assert isinstance(reward, np.float64) and isinstance(done, np.bool_)
        assert set(info) == {'is_success', 'is_crash', 'distance'}
        assert info['is_success'].dtype == np.float32 and info['distance'].dtype == np.float64
        o = obs['observation']
        assert np.array_equal(o[512:514], prev['achieved_goal'])     # prev_pose
        assert np.array_equal(o[514:516], obs['achieved_goal'])      # pose
        if t > 0:
            assert np.allclose(o[516:518], prev_a)                   # vel = previous action
        assert env.steps_since_reset == t + 1
        # the HER entry points agree with what step returned (float32 observation round trip)
        if not info['is_crash']:
            assert abs(env.compute_reward(a, obs) - reward) < 1e-4
            assert env.compute_done(obs) == done
            ci = env.compute_info(obs)
            assert ci['is_success'] == info['is_success'] and abs(ci['distance'] - info['distance']) < 1e-5
        prev, prev_a = obs, a
        if done:
            break
    assert abs(env.robot.px - obs['achieved_goal'][0]) < 1e-12
    # the batch's host export carries the same attribute surface, and the
    # debug view draws from either
    view = env._sim.export_env(0)
    assert view.map_info is env.map_info and view.num_scan_stack == env.num_scan_stack
    assert (view.robot.px, view.robot.py, view.robot.theta) == (env.robot.px, env.robot.py, env.robot.theta)
    assert (view.robot.gx, view.robot.gy) == (env.robot.gx, env.robot.gy)
    assert len(view.humans) == len(env.humans)
    assert all((a.px, a.py, a.theta) == (b.px, b.py, b.theta) for a, b in zip(view.humans, env.humans))
    assert np.array_equal(view.prev_obs['observation'], env.prev_obs['observation'])
    assert view.steps_since_reset == env.steps_since_reset
    img = env.render(mode='rgb_array')
    assert img.shape == (800, 800, 3) and img.dtype == np.uint8
    from nav_gym_b200.render import render
    assert np.array_equal(render(view, 'rgb_array'), img)


def test_pedestrian_kernel():
    from nav_gym_b200 import maps, _lib
    from nav_gym_b200.batched_env import BatchedNavGym
    from nav_gym_b200.robot import legs_to_discs, footprint_segments, Human
    rng = np.random.RandomState(0)
    m = maps.create_outdoor_map(10, 0.7, rng)
    B, P = 8, 5
    pool = maps.spawn_pool(m, 64, rng, min_goal_dist=5, max_goal_dist=15)
    rows = pool[:B]
    peds = np.stack([maps.spawn_pedestrians(m, rows[e, :2], P, rng) for e in range(B)])
    env = BatchedNavGym(B, [m], device='cuda:0')
    env.set_state(rows[:, :2], rows[:, 2:4], rows[:, 4])
    env.attach_pedestrians(peds)
    env.reset()
    nd, ns = env._pnd.cpu().numpy(), env._pns.cpu().numpy()
    legs = peds[:, :, 12] > 0.5
    assert np.array_equal(nd, 2 * legs.sum(1)) and np.array_equal(ns, 4 * (~legs).sum(1))
    # geometry at the initial poses == the host helpers
    d = env._pdiscs.cpu().numpy()
    s = env._psegs.cpu().numpy()
    for e in range(B):
        want_d = [legs_to_discs(peds[e, p, :3], peds[e, p, 9:12]) for p in range(P) if legs[e, p]]
        if want_d:
            assert np.allclose(d[e, :nd[e]], np.concatenate(want_d), atol=1e-5)
        want_s = [footprint_segments(*peds[e, p, :3], Human.footprint) for p in range(P) if not legs[e, p]]
        if want_s:
            assert np.allclose(s[e, :ns[e]], np.concatenate(want_s), atol=1e-5)
    # motion: speed * dt per step towards the target, unicycle update of human.py:32-41
    act = torch.zeros(B, 2, device='cuda')
    p0 = env.peds.cpu().numpy().copy()
    for _ in range(60):
        env.step(act)
    p1 = env.peds.cpu().numpy()
    moved = np.hypot(p1[:, :, 0] - p0[:, :, 0], p1[:, :, 1] - p0[:, :, 1])
    assert np.all(moved <= p0[:, :, 3] * 0.2 * 60 + 1e-3)
    far = np.hypot(p0[:, :, 6] - p0[:, :, 0], p0[:, :, 7] - p0[:, :, 1]) > 9
    d_before = np.hypot(p0[:, :, 6] - p0[:, :, 0], p0[:, :, 7] - p0[:, :, 1])
    d_after = np.hypot(p1[:, :, 6] - p1[:, :, 0], p1[:, :, 7] - p1[:, :, 1])
    fast = p0[:, :, 3] > 0.2
    assert np.all(d_after[far & fast] < d_before[far & fast])   # they approach their goals
    assert np.all(p1[:, :, 9:12] != p0[:, :, 9:12]) or not fast.any()
    # a pedestrian parked right in front of the robot shows up in the scan
    peds2 = peds.copy()
    peds2[:, 0, 0] = rows[:, 0] + 1.5 * np.cos(rows[:, 4])
    peds2[:, 0, 1] = rows[:, 1] + 1.5 * np.sin(rows[:, 4])
    peds2[:, 0, 3] = 0
    peds2[:, 0, 4:6] = peds2[:, 0, 0:2]
    peds2[:, 0, 6:8] = peds2[:, 0, 0:2]
    peds2[:, 0, 12] = 0   # box
    env2 = BatchedNavGym(B, [m], device='cuda:0')
    env2.set_state(rows[:, :2], rows[:, 2:4], rows[:, 4])
    env2.attach_pedestrians(peds2)
    obs = env2.reset().cpu().numpy()
    assert np.all(obs[:, 256] < 1.5) and np.all(obs[:, 256] > 1.2)   # beam 256 looks forward


def test_gym_make_scan_stack():
    """num_scan_stack = 3 through the drop-in: observation = [3 x 512 scans | 7], newest last,
    missing history padded with the current scan (env.py:257-279)."""
    import nav_gym_b200  # noqa: F401
    from nav_gym_b200 import gym_shim
    gym = gym_shim.install()
    np.random.seed(5)
    env = gym.make('NavGym-v0', num_scan_stack=3)
    assert env.observation_space['observation'].shape == (3 * 512 + 7,)
    o0 = env.reset()['observation']
    assert o0.shape == (1543,)
    s0 = o0[1024:1536]
    assert np.array_equal(o0[:512], s0) and np.array_equal(o0[512:1024], s0)
    o1 = env.step(np.array([0.3, 0.1]))[0]['observation']
    s1 = o1[1024:1536]
    assert np.array_equal(o1[:512], s1) and np.array_equal(o1[512:1024], s0)
    o2, r, d, info = env.step(np.array([0.3, -0.1]))
    o2 = o2['observation']
    if not info['is_crash']:
        assert np.array_equal(o2[:512], s0) and np.array_equal(o2[512:1024], s1)


def test_her_kernel_known_answers_batch():
    """6000 stored observations (clear / in discomfort / crashed / at the goal; more rows than one
    wave of the kernel's warps, so every warp strides over several) against the outputs of the
    reference's own compute_rewards / compute_terminals on the same float64 rows
    (tests/golden/her_batch.npz, minted by oracle/make_golden_her.py)."""
    import os
    import synth
    from nav_gym_b200.batched_env import BatchedNavGym
    G = np.load(os.path.join(gu.GOLDEN, 'her_batch.npz'))
    rng = np.random.RandomState(0)
    env = BatchedNavGym(1, [synth.outdoor_map(rng, size=100, n_obs=2)], device='cuda:0')
    assert np.array_equal(env.scan_threshold, G['thr']) and np.array_equal(env.scan_discomfort_threshold, G['dthr'])
    b = synth.her_batch()
    rows = synth.her_rows(b, G['thr'], G['dthr'])
    out = env.compute_rewards(torch.from_numpy(rows).cuda(), torch.from_numpy(b['goal']).cuda())
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert np.array_equal(out['done'].astype(bool), G['done'])
    # float32 output of the float64 sum the reference forms: half an ulp at |reward| <= 16
    assert np.allclose(out['reward'], G['reward'], rtol=0, atol=1.0e-6)
    assert (b['kind'] == 1).sum() > 1000 and (out['is_crash'] == 1).sum() > 1000
    # a strided, padded view of the same rows (obs_stride > 519) gives the same answers
    wide = torch.zeros(len(rows), 600, device='cuda')
    wide[:, :519] = torch.from_numpy(rows).cuda()
    out2 = env.compute_rewards(wide[:, :519], torch.from_numpy(b['goal']).cuda())
    assert torch.equal(out2['reward'].cpu(), torch.from_numpy(out['reward']))
