"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
and the golden fixtures.  Integer / flag outputs and float32 scans bit-exact; float64 poses
within 1e-9 m, rewards within 1e-6 (float32 output of a float64 sum)."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_util as gu
import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _maps(seed=0):
    rng = np.random.RandomState(seed)
    return [synth.indoor_map(rng), synth.outdoor_map(rng), synth.outdoor_map(rng, size=300, n_obs=25)]


def test_thresholds_match_oracle():
    from nav_gym_b200.batched_env import scan_thresholds
    thr, dthr = scan_thresholds('cuda:0')
    assert np.array_equal(thr, orc.footprint_threshold(orc.THRESHOLD_FOOTPRINT))
    assert np.array_equal(dthr, orc.footprint_threshold(orc.DISCOMFORT_FOOTPRINT))


@pytest.mark.parametrize('name', gu.trace_names()[:2] + ['synthetic'])
def test_edt_exact(name):
    """Device EDT (integer brute-force row pass) == oracle EDT (Felzenszwalb-Huttenlocher)."""
    from nav_gym_b200.batched_env import MapPool
    from nav_gym_b200 import natives
    maps = _maps() if name == 'synthetic' else [gu.map_info(gu.load(name))]
    empty = dict(data=np.zeros((64, 96), np.int8), origin=(0, 0), resolution=0.05, width=96, height=64)
    one = dict(data=np.zeros((50, 40), np.int8), origin=(0, 0), resolution=0.05, width=40, height=50)
    one['data'][17, 3] = 100
    maps = maps + [empty, one]
    pool = MapPool(maps, 'cuda:0')
    for i, m in enumerate(maps):
        want = orc.edt(np.asarray(m['data']) >= 0.1)
        assert np.array_equal(pool.edt(i).cpu().numpy(), want), 'map %d' % i
    # the same transform behind the range_libc drop-in
    rm = natives.PyRayMarching(natives.PyOMap(np.asarray(maps[0]['data']) >= 0.1), 1e6)
    assert np.array_equal(rm.edt_host(), orc.edt(np.asarray(maps[0]['data']) >= 0.1))


def test_calc_range_many_bit_exact():
    from nav_gym_b200 import natives
    rng = np.random.RandomState(3)
    for m in _maps(1):
        occ = np.ascontiguousarray(np.asarray(m['data']) >= 0.1)
        dist = orc.edt(occ)
        H, W = occ.shape
        n = 20000
        ins = np.column_stack([rng.randint(0, W, n), rng.randint(0, H, n),
                               rng.uniform(-7, 7, n)]).astype(np.float32)
        rm = natives.PyRayMarching(natives.PyOMap(occ), W * H)
        outs = np.zeros(n, np.float32)
        rm.calc_range_many(ins, outs)
        want, hits = orc.calc_range_many(dist, ins, float(W * H), want_hits=True)
        assert np.array_equal(outs, want)
        got_r, got_h = rm.calc_range_many_with_hits(ins)
        assert np.array_equal(got_r, want) and np.array_equal(got_h, hits.astype(np.int16))


@pytest.mark.parametrize('early_stop', [False, True])
@pytest.mark.parametrize('name', gu.trace_names())
def test_cuda_replays_reference_trace(name, early_stop):
    import cuda_util
    G = gu.load(name)
    st = cuda_util.golden_stepper(G, early_stop)
    assert np.array_equal(st.env.scan_threshold, G['thr'])
    assert np.array_equal(st.env.scan_discomfort_threshold, G['dthr'])
    gu.replay(G, st, pose_tol=1e-9, reward_tol=2e-6)


@pytest.mark.parametrize('B,T,seed', [(256, 40, 0), (1024, 12, 1)])
def test_batched_step_matches_oracle(B, T, seed):
    """Many envs over several maps, pedestrians as discs + box segments, injected noise,
    random actions: every step compared with the oracle's lockstep batch."""
    import cuda_util
    rng = np.random.RandomState(seed)
    maps = _maps(seed + 10)
    map_id = rng.randint(0, len(maps), B).astype(np.int32)
    edts = [orc.edt(np.asarray(m['data']) >= 0.1) for m in maps]
    start = np.zeros((B, 2))
    for i, m in enumerate(maps):
        sel = np.where(map_id == i)[0]
        start[sel] = synth.free_poses(rng, m, len(sel), 14, edts[i])
    goal = start + rng.uniform(-3, 3, (B, 2))
    theta = rng.uniform(0, 2 * np.pi, B)
    md, ms = 12, 16
    o = orc.OracleBatch(maps, map_id, start, goal, theta, params=dict(t_stop=502.0), max_disc=md, max_seg=ms)
    c = cuda_util.CudaStepper(maps, map_id, start, goal, theta, max_disc=md, max_seg=ms, early_stop=True)
    geom = synth.random_geometry(rng, B, start, md, ms)
    noise = (0.02 * rng.randn(B, 2, 512)).astype(np.float32)
    o.reset_obs(*geom, noise=noise)
    c.reset_obs(*geom, noise=noise)
    _compare(o, c, 'reset', first=True)
    n_crash = n_succ = n_disc = 0
    for t in range(T):
        act = rng.uniform([-0.1, -0.7], [0.55, 0.7], (B, 2)).astype(np.float32)
        pos = o.state[:2].T.copy()
        geom = synth.random_geometry(rng, B, pos, md, ms)
        noise = (0.02 * rng.randn(B, 2, 512)).astype(np.float32)
        o.step(act, *geom, noise=noise)
        c.step(act, *geom, noise=noise)
        _compare(o, c, 'step %d' % t)
        n_crash += int(o.is_crash.sum())
        n_succ += int(o.is_success.sum())
    assert n_crash > 0 and n_succ > 0


def test_obstacle_lists_at_capacity():
    """The obstacle phase lists the visible obstacles in 256 shared-memory words (segments from the
    front, discs from the back): 128 segments + 64 discs per environment, every list full and
    everything close to the robot (nothing out of sight, some obstacles touching it: whole-scan
    windows), some environments with far-away obstacles only (empty lists) -- scans bit-exact
    against the oracle's all-beams loop."""
    import cuda_util
    rng = np.random.RandomState(21)
    B, md, ms = 96, 64, 128
    maps = _maps(31)
    map_id = rng.randint(0, len(maps), B).astype(np.int32)
    edts = [orc.edt(np.asarray(m['data']) >= 0.1) for m in maps]
    start = np.zeros((B, 2))
    for i, m in enumerate(maps):
        sel = np.where(map_id == i)[0]
        start[sel] = synth.free_poses(rng, m, len(sel), 14, edts[i])
    goal = start + rng.uniform(-3, 3, (B, 2))
    theta = rng.uniform(0, 2 * np.pi, B)
    o = orc.OracleBatch(maps, map_id, start, goal, theta, params=dict(t_stop=502.0), max_disc=md, max_seg=ms)
    c = cuda_util.CudaStepper(maps, map_id, start, goal, theta, max_disc=md, max_seg=ms, early_stop=True)

    def geometry(pos):
        discs, ndisc, segs, nseg = synth.random_geometry(rng, B, pos, md, ms, spread=3.0)
        ndisc[:], nseg[:] = md, ms                      # full lists (unused rows of segs are filled below)
        for e in range(B):
            for b in range(ms // 4):
                if not segs[e, 4 * b:4 * b + 4].any():   # a copy of the first box, shifted
                    segs[e, 4 * b:4 * b + 4] = segs[e, 0:4] + np.tile(rng.uniform(-3, 3, 2).astype(np.float32), 2)
        discs[:8, 0, :2] = pos[:8] + 0.01                # a disc on top of the robot
        segs[8:16, 0] = np.concatenate([pos[8:16] - 0.5, pos[8:16] + 0.5], axis=1)   # a segment through it
        far = np.arange(B) % 7 == 3                      # environments that see nothing
        discs[far, :, :2] += 400.0
        segs[far] += 400.0
        return discs, ndisc, segs, nseg

    noise = np.zeros((B, 2, 512), np.float32)
    g = geometry(start)
    o.reset_obs(*g, noise=noise)
    c.reset_obs(*g, noise=noise)
    _compare(o, c, 'reset', first=True)
    hit_any = 0
    for t in range(6):
        act = rng.uniform([0.0, -0.64], [0.5, 0.64], (B, 2)).astype(np.float32)
        g = geometry(o.state[:2].T.copy())
        o.step(act, *g, noise=noise)
        c.step(act, *g, noise=noise)
        _compare(o, c, 'step %d' % t)
        hit_any += int((o.obs[:, :512] < 3.0).sum())
    assert hit_any > 10000


def _compare(o, c, tag, first=False):
    assert np.array_equal(c.hits, o.hits), tag + ' hit cells'
    assert np.array_equal(c.obs[:, :512], o.obs[:, :512]), tag + ' scan'
    assert np.array_equal(c.steps, o.steps), tag
    assert np.allclose(c.state, o.state, rtol=0, atol=1e-9), tag + ' state'
    assert np.allclose(c.tail64, o.tail64, rtol=0, atol=1e-9), tag + ' tail'
    assert np.allclose(c.obs[:, 512:], o.obs[:, 512:], rtol=1e-6, atol=1e-6), tag
    if not first:
        assert np.array_equal(c.done, o.done), tag + ' done'
        assert np.array_equal(c.is_crash, o.is_crash), tag + ' crash'
        assert np.array_equal(c.is_success, o.is_success), tag + ' success'
        assert np.allclose(c.reward, o.reward, rtol=1e-6, atol=2e-6), tag + ' reward'
        assert np.allclose(c.distance, o.distance, rtol=1e-6, atol=1e-5), tag


def test_abi_host_render_matches_oracle():
    from nav_gym_b200 import natives
    rng = np.random.RandomState(5)
    K = 512
    angles = np.linspace(-3.141592, 3.141592 - 0.0122718463, K) + np.float32(1.234)
    for _ in range(5):
        lidar = rng.uniform(2, 8, 2).astype(np.float32)
        contours = [(lidar + rng.uniform(-4, 4, 2) + np.array([[0.3, 0.2], [-0.3, 0.2], [-0.3, -0.2], [0.3, -0.2]])).tolist()
                    for _ in range(6)]
        r1 = np.full(K, 25, np.float32)
        natives.render_contours_in_lidar(r1, angles, natives.flatten_contours(contours), lidar)
        r2 = np.full(K, 25, np.float32)
        _, dirs = orc.beam_dirs(np.float32(0), lin=angles.astype(np.float32).astype(np.float64))
        orc.render_contours(r2, dirs, orc.flatten_contours(contours), lidar)
        assert np.array_equal(r1, r2)
        assert (r1 < 25).any()


def test_step_host_pipeline_matches_device_step():
    """The chunked host-buffer call (prioritised streams, overlapped D2H) returns exactly what
    the device-resident step computes, including auto-reset and in-kernel Philox noise."""
    from nav_gym_b200 import maps
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool
    rng = np.random.RandomState(4)
    m = maps.create_outdoor_map(10, 0.7, rng)
    pool = filter_spawn_pool(m, maps.spawn_pool(m, 2048, rng, min_goal_dist=4, max_goal_dist=15))
    mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
    B = 1000
    envs = [BatchedNavGym(B, mp, seed=9, auto_reset=True) for _ in range(2)]
    for e in envs:
        e.reset_from_spawn_pool(np.random.RandomState(5))
    obs_h = torch.empty(B, 519).pin_memory()
    rew_h = torch.empty(B).pin_memory()
    done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    n_done = 0
    for t in range(30):
        act = torch.from_numpy(rng.uniform([0.2, -0.64], [0.5, 0.64], (B, 2)).astype(np.float32))
        o, r, d, _ = envs[0].step(act.cuda())
        torch.cuda.synchronize()
        envs[1].step_host(act.pin_memory(), obs_h, rew_h, done_h, chunks=3)
        assert torch.equal(o.cpu(), obs_h) and torch.equal(r.cpu(), rew_h) and torch.equal(d.cpu(), done_h), t
        n_done += int(d.sum())
    assert n_done > 0
    assert torch.equal(envs[0].state, envs[1].state) and torch.equal(envs[0].episodes, envs[1].episodes)


def test_async_host_groups_match_device_step():
    """Two env groups kept in flight through submit_host / wait_host give, group by group,
    exactly the device-resident step's results."""
    from nav_gym_b200 import maps
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool
    rng = np.random.RandomState(6)
    m = maps.create_outdoor_map(10, 0.7, rng)
    pool = filter_spawn_pool(m, maps.spawn_pool(m, 2048, rng, min_goal_dist=4, max_goal_dist=15))
    mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
    B = 777
    envs = [BatchedNavGym(B, mp, seed=3, auto_reset=True) for _ in range(2)]
    for e in envs:
        e.reset_from_spawn_pool(np.random.RandomState(7))
    act_h = torch.empty(B, 2).pin_memory()
    obs_h = torch.empty(B, 519).pin_memory()
    rew_h = torch.empty(B).pin_memory()
    done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    bounds = envs[1].host_groups(2, act_h, obs_h, rew_h, done_h)
    T = 12
    acts = torch.from_numpy(rng.uniform([0.2, -0.64], [0.5, 0.64], (T, B, 2)).astype(np.float32))
    want = []
    for t in range(T):
        o, r, d, _ = envs[0].step(acts[t].cuda())
        want.append((o.cpu().clone(), r.cpu().clone(), d.cpu().clone()))
    act_h.copy_(acts[0])
    for g in range(2):
        envs[1].submit_host(g)
    for t in range(T):
        for g, (b0, b1) in enumerate(bounds):
            envs[1].wait_host(g)
            assert torch.equal(obs_h[b0:b1], want[t][0][b0:b1]), (t, g)
            assert torch.equal(rew_h[b0:b1], want[t][1][b0:b1]) and torch.equal(done_h[b0:b1], want[t][2][b0:b1])
            if t + 1 < T:
                act_h[b0:b1].copy_(acts[t + 1][b0:b1])
                envs[1].submit_host(g)


def test_map_pool_auto_reset_with_pedestrians_is_deterministic():
    """C4-style world: several maps, per-env pedestrian counts, map re-drawn at auto-reset.
    Invariants + run-to-run determinism (the Philox streams are keyed by env / episode / step)."""
    from nav_gym_b200 import maps, _lib
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool
    rng = np.random.RandomState(8)
    ms = [maps.create_outdoor_map(10, 0.5, rng), maps.create_indoor_map(3, 40, rng, cells=40),
          maps.create_outdoor_map(10, 0.9, rng, size=300)]
    pools = [filter_spawn_pool(m, maps.spawn_pool(m, 512, rng, min_goal_dist=3, max_goal_dist=12)) for m in ms]
    assert all(len(p) > 20 for p in pools)
    mp = MapPool(ms, 'cuda:0', spawn_pools=pools)
    B, P = 300, 6
    map_id = rng.randint(0, 3, B).astype(np.int32)
    peds = np.stack([maps.spawn_pedestrians(ms[map_id[e]], (-50, -50), P, np.random.RandomState(e)) for e in range(B)])
    nped = rng.randint(0, P + 1, B).astype(np.int32)
    acts = torch.from_numpy(rng.uniform([0.2, -0.64], [0.5, 0.64], (40, B, 2)).astype(np.float32)).cuda()
    runs = []
    for rep in range(2):
        env = BatchedNavGym(B, mp, map_id=map_id, seed=11, auto_reset=True, resample_map=True,
                            max_episode_steps=25)
        env.reset_from_spawn_pool(np.random.RandomState(9))
        env.attach_pedestrians(peds, nped=nped)
        env.reset()
        log = []
        for t in range(40):
            o, r, d, info = env.step(acts[t])
            log.append((o.clone(), r.clone(), d.clone(), env.map_id.clone(), info['truncated'].clone()))
        runs.append(log)
        assert int(env.episodes.sum()) > B // 4                      # episodes ended and restarted
        assert int(sum(x[4].sum() for x in log)) > 0                  # some by the step limit
        mid = env.map_id.cpu().numpy()
        assert mid.min() >= 0 and mid.max() <= 2 and len(np.unique(mid)) == 3
        o = log[-1][0].cpu().numpy()
        assert np.isfinite(o).all() and o[:, :512].min() > -0.5 and o[:, :512].max() < 25.5
        assert int(env.steps.max()) <= 25
    for a, b in zip(*runs):
        assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_no_fma_build_matches_no_fma_oracle():
    """The -DNAVGYM_MARCH_NO_FMA library (x0 + dx * t rounded in two steps) against the oracle
    built the same way: the pair the real range_libc would be matched with if its binary does
    not contract the multiply-add.  Runs in a subprocess (one process loads one library)."""
    import os
    import subprocess
    import sys
    from nav_gym_b200 import _lib
    if not os.path.exists(_lib.SO_NOFMA):
        pytest.skip('libnavgym_b200_nofma.so not built (python -c "import __graft_entry__ as g; g.build()")')
    code = r'''
import numpy as np, sys, os
sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import synth, cuda_util
from oracle import oracle as orc
from nav_gym_b200 import _lib, natives
assert _lib.load().navgym_march_is_fused() == 0 and orc.lib().nvo_march_is_fused() == 0
rng = np.random.RandomState(2)
maps = [synth.indoor_map(rng), synth.outdoor_map(rng)]
for m in maps:
    occ = np.ascontiguousarray(np.asarray(m['data']) >= 0.1)
    dist = orc.edt(occ)
    H, W = occ.shape
    n = 400000
    ins = np.column_stack([rng.randint(0, W, n), rng.randint(0, H, n), rng.uniform(-7, 7, n)]).astype(np.float32)
    rm = natives.PyRayMarching(natives.PyOMap(occ), W * H)
    got, hits = rm.calc_range_many_with_hits(ins)
    want, wh = orc.calc_range_many(dist, ins, float(W * H), want_hits=True)
    assert np.array_equal(got, want) and np.array_equal(hits, wh.astype(np.int16))
B = 512
map_id = rng.randint(0, 2, B).astype(np.int32)
edts = [orc.edt(np.asarray(m['data']) >= 0.1) for m in maps]
start = np.zeros((B, 2))
for i, m in enumerate(maps):
    sel = np.where(map_id == i)[0]
    start[sel] = synth.free_poses(rng, m, len(sel), 14, edts[i])
goal = start + rng.uniform(-3, 3, (B, 2)); theta = rng.uniform(0, 2 * np.pi, B)
o = orc.OracleBatch(maps, map_id, start, goal, theta, params=dict(t_stop=502.0))
c = cuda_util.CudaStepper(maps, map_id, start, goal, theta, early_stop=True)
o.reset_obs(); c.reset_obs()
for t in range(10):
    act = rng.uniform([-0.1, -0.7], [0.55, 0.7], (B, 2)).astype(np.float32)
    o.step(act); c.step(act)
    assert np.array_equal(c.hits, o.hits) and np.array_equal(c.obs[:, :512], o.obs[:, :512]), t
    assert np.array_equal(c.done, o.done)
print('no-fma pair ok')
'''
    env = dict(os.environ, NAVGYM_LIB=_lib.SO_NOFMA, NAVGYM_ORACLE_VARIANT='_nofma')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'no-fma pair ok' in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_level1_native_call_transcript():
    """INTEGRATION.md level 1, end to end on the device: every call the unmodified reference
    env.py made across its native boundary during one episode (tests/golden/native_calls.npz,
    minted by oracle/make_golden_native_calls.py -- PyOMap / PyRayMarching construction, then per
    _compute_scan of the robot and of every pedestrian: calc_range_many, render_contours_in_lidar,
    CMap2D.render_agents_in_lidar with CSimAgent legs) is replayed, in order and with the
    reference's own arguments, through nav_gym_b200.natives; every result equals the transcript
    bit for bit.  (The transcript's results come from the oracle-backed stand-ins: the native
    semantics themselves stay "parity unpinned".)"""
    from nav_gym_b200 import natives
    G = gu.load('native_calls')
    names = [str(x) for x in G['kind_names']]
    rm, n = None, {k: 0 for k in names}
    for i, kind in enumerate(G['kind']):
        fn = names[int(kind)]
        g = lambda k: G['c%d_%s' % (i, k)]
        if fn == 'PyRayMarching':
            H, W = [int(v) for v in g('occ_shape')]
            occ = np.ascontiguousarray(np.unpackbits(g('occ_bits'))[:H * W].reshape(H, W).astype(np.bool_))
            rm = natives.PyRayMarching(natives.PyOMap(occ), float(g('max_range')))   # env.py:337-340
        elif fn == 'calc_range_many':
            ins = np.ascontiguousarray(g('ins'))
            outs = np.zeros(len(ins), np.float32)
            rm.calc_range_many(ins, outs)                                            # env.py:425
            assert np.array_equal(outs, g('outs')), 'call %d calc_range_many' % i
        elif fn == 'render_contours_in_lidar':
            r = g('ranges_in').copy()
            natives.render_contours_in_lidar(r, g('angles'), g('flat'), g('lidar_xy'))   # env.py:430-431
            assert np.array_equal(r, g('ranges_out')), 'call %d render_contours_in_lidar' % i
        else:
            r = g('ranges_in').copy()
            agents = [natives.CSimAgent(p, s_, v) for p, s_, v in zip(g('poses'), g('states'), g('vels'))]
            cm = natives.CMap2D()
            cm.render_agents_in_lidar(r, g('angles'), agents, g('lidar_xy'))         # env.py:432
            assert np.array_equal(r, g('ranges_out')), 'call %d render_agents_in_lidar' % i
        n[fn] += 1
    assert n['PyRayMarching'] == 1 and min(n.values()) >= 1 and sum(n.values()) == len(G['kind']) > 60
