"""Small run of every kernel of the library, for compute-sanitizer (tools/sanitize.sh):
the fused step with scripted pedestrians, Philox noise, auto-reset over a map pool and a scan
stack; the host-buffer pipe; the HER kernel; the stand-alone natives; one policy-driven crowd
step (pedestrian lidar, routes, motion, the three tcgen05 policy launches).  Sizes are tiny:
the sanitizer tools slow a launch down 10-100x.  `python tools/sanitize_target.py [part ...]`
with parts robot, host, natives, crowd (default: all)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nav_gym_b200 import maps, _lib, natives
from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool

dev = 'cuda:0'
rng = np.random.RandomState(3)
ms = [maps.create_outdoor_map(10, 0.5, rng), maps.create_indoor_map(3, 40, rng, cells=40)]
pools = [filter_spawn_pool(m, maps.spawn_pool(m, 256, rng, min_goal_dist=3, max_goal_dist=12)) for m in ms]
mp = MapPool(ms, dev, spawn_pools=pools)


def acts(B, T, seed):
    r = np.random.RandomState(seed)
    return torch.from_numpy(r.uniform([0.2, -0.64], [0.5, 0.64], (T, B, 2)).astype(np.float32)).to(dev)


def robot():
    B, P, T = 96, 6, 12
    map_id = rng.randint(0, 2, B).astype(np.int32)
    peds = np.stack([maps.spawn_pedestrians(ms[map_id[e]], (-50, -50), P, np.random.RandomState(e)) for e in range(B)])
    nped = rng.randint(0, P + 1, B).astype(np.int32)
    a = acts(B, T, 1)
    for stack, trunk in ((1, False), (3, True)):
        env = BatchedNavGym(B, mp, map_id=map_id, seed=11, auto_reset=True, resample_map=True,
                            max_episode_steps=5, num_scan_stack=stack, record_hits=True)
        env.reset_from_spawn_pool(np.random.RandomState(9))
        env.attach_pedestrians(peds, nped=nped, trunk_mode=trunk)
        env.reset()
        for t in range(T):
            env.step(a[t])
        torch.cuda.synchronize()
        assert torch.isfinite(env.obs).all() and int(env.episodes.sum()) > 0
        if stack == 1:
            env.compute_rewards(env.obs, env.state[3:5].T.contiguous().float())
        env.export_env(0)
    torch.cuda.synchronize()
    print('robot ok')


def host():
    B, T = 128, 6
    env = BatchedNavGym(B, mp, map_id=rng.randint(0, 2, B).astype(np.int32), seed=5, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(2))
    a = acts(B, T, 2).cpu()
    act_h = torch.zeros(B, 2).pin_memory()
    obs_h = torch.zeros(B, env.obs_dim).pin_memory()
    rew_h = torch.zeros(B).pin_memory()
    done_h = torch.zeros(B, dtype=torch.uint8).pin_memory()
    for t in range(T):
        act_h.copy_(a[t])
        env.step_host(act_h, obs_h, rew_h, done_h, chunks=2)
    bounds = env.host_groups(2, act_h, obs_h, rew_h, done_h)
    for g in range(2):
        env.submit_host(g)
    for t in range(T):
        for g in range(2):
            env.wait_host(g)
            if t + 1 < T:
                env.submit_host(g)
    assert torch.isfinite(obs_h).all()
    print('host ok', bounds)


def native_calls():
    m = ms[0]
    occ = np.ascontiguousarray(np.asarray(m['data']) >= 0.1)
    om = natives.PyOMap(occ)
    rm = natives.PyRayMarching(om, float(occ.size))
    ins = np.zeros((512, 3), np.float32)
    ins[:, 0], ins[:, 1] = 200, 200
    ins[:, 2] = np.linspace(-3.14, 3.14, 512)
    outs = np.zeros(512, np.float32)
    rm.calc_range_many(ins, outs)
    ranges = (outs * 0.05).astype(np.float32)
    angles = ins[:, 2].copy()
    flat = natives.flatten_contours([[[11.0, 10.0], [11.5, 10.0], [11.5, 10.5], [11.0, 10.5]]])
    natives.render_contours_in_lidar(ranges, angles, flat, np.array([10.0, 10.0], np.float32))
    cm = natives.CMap2D()
    cm.set_resolution(1.)
    ag = natives.CSimAgent(np.array([9.0, 10.0, 0.3], np.float32), np.array([0.4, 0.1, 0.2], np.float32), np.zeros(2, np.float32))
    cm.render_agents_in_lidar(ranges, angles, [ag], np.array([10.0, 10.0], np.float32))
    assert np.isfinite(ranges).all()
    print('natives ok')


def crowd():
    from nav_gym_b200.pedestrians import PedestrianSim
    B, P, T = 160, 3, 3          # 480 pedestrians: three full 128-row policy tiles and a ragged one
    env = BatchedNavGym(B, mp, map_id=rng.randint(0, 2, B).astype(np.int32), seed=2, auto_reset=True, max_episode_steps=2)
    env.reset_from_spawn_pool(np.random.RandomState(2))
    sim = PedestrianSim(env, P, seed=2)
    a = acts(B, T, 3)
    for t in range(T):
        sim.step(a[t])
    torch.cuda.synchronize()
    assert torch.isfinite(sim.pose).all() and torch.isfinite(sim.scan).all() and torch.isfinite(env.obs).all()
    print('crowd ok')


if __name__ == '__main__':
    parts = dict(robot=robot, host=host, natives=native_calls, crowd=crowd)
    for p in (sys.argv[1:] or list(parts)):
        parts[p]()
