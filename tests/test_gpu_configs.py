"""GPU: parity at the parameters of BASELINE.json's configs[2] (C3: 16 384 envs on a 2000 x 2000
outdoor map, 20 pedestrians per env = up to 40 discs / 80 segments) and configs[3] (C4: one GPU's
8192-env share of the 16-map pool, 5..15 pedestrians, map re-drawn at auto-reset), the Philox
production noise as a distribution, the C rollout driver, and a non-square map.

The full batches run on the device with device-scripted pedestrians and in-kernel auto-reset; a
random subsample of environments is shadowed by the CPU oracle: before every step the shadow
takes the device's state of those environments, then steps them with the same actions, the
geometry the device's pedestrian kernel emitted and the same injected noise, and every output is
compared (hit cells, scans, flags bit-exact; poses 1e-9; rewards 2e-6).  Environments that ended
their episode are auto-reset on the device: their returned observation is compared with the
oracle's first observation of the state the device drew, and that state must be a row of the new
map's spawn pool."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
NB = 512


class Shadow(object):
    """CPU oracle shadow of the environments `pick` of a BatchedNavGym with scripted pedestrians."""

    def __init__(self, env, maps, pick):
        self.env, self.pick = env, pick
        n = len(pick)
        z2 = np.zeros((n, 2))
        kw = dict(params=dict(t_stop=502.0), max_disc=env.max_disc, max_seg=env.max_seg)
        self.o = orc.OracleBatch(maps, np.zeros(n, np.int32), z2, z2, np.zeros(n), **kw)
        self.o2 = orc.OracleBatch(maps, np.zeros(n, np.int32), z2, z2, np.zeros(n), **kw)
        self.pools = env.pool.spawn_arrays

    def _take(self, o):
        env, pick = self.env, self.pick
        o.state[:] = env.state[:, pick].cpu().numpy()
        o.steps[:] = env.steps[pick].cpu().numpy()
        o.map_id[:] = env.map_id[pick].cpu().numpy()

    def _geom(self):
        env, pick = self.env, self.pick
        return (env._pdiscs[pick].cpu().numpy(), env._pnd[pick].cpu().numpy(),
                env._psegs[pick].cpu().numpy(), env._pns[pick].cpu().numpy())

    def step(self, act, noise, tag):
        """act [B, 2] / noise [B, 2, 512] device tensors.  Returns (#done, #crash) of the sample."""
        env, pick, o, o2 = self.env, self.pick, self.o, self.o2
        torch.cuda.synchronize()
        self._take(o)
        env.step(act, noise=noise)
        torch.cuda.synchronize()
        geom = self._geom()
        nz = noise[pick].cpu().numpy()
        o.step(act[pick].cpu().numpy(), *geom, noise=nz)
        g = lambda name: getattr(env, name)[pick].cpu().numpy()
        assert np.array_equal(g('hits'), o.hits), tag + ' hit cells'
        assert np.array_equal(g('done'), o.done), tag + ' done'
        assert np.array_equal(g('is_crash'), o.is_crash), tag + ' crash'
        assert np.array_equal(g('is_success'), o.is_success), tag + ' success'
        assert np.allclose(g('reward'), o.reward, rtol=1e-6, atol=2e-6), tag + ' reward'
        assert np.allclose(g('distance'), o.distance, rtol=1e-6, atol=1e-5), tag + ' distance'
        done = o.done.astype(bool)
        reset = done if env.args.auto_reset else np.zeros_like(done)
        keep = ~reset
        obs, tail, st = g('obs'), g('tail64'), env.state[:, pick].cpu().numpy()
        assert np.array_equal(obs[keep, :NB], o.obs[keep, :NB]), tag + ' scan'
        assert np.allclose(tail[keep], o.tail64[keep], rtol=0, atol=1e-9), tag + ' tail'
        assert np.allclose(st[:, keep], o.state[:, keep], rtol=0, atol=1e-9), tag + ' state'
        assert np.array_equal(g('steps')[keep], o.steps[keep]), tag + ' steps'
        if reset.any():
            # the device drew a new episode: its first observation, against the oracle's
            self._take(o2)
            o2.reset_obs(*geom, noise=nz)
            assert np.array_equal(obs[reset, :NB], o2.obs[reset, :NB]), tag + ' first scan after auto-reset'
            assert np.allclose(tail[reset], o2.tail64[reset], rtol=0, atol=1e-9), tag
            assert (g('steps')[reset] == 0).all()
            mid = o2.map_id
            for i in np.where(reset)[0]:
                row = st[[orc.S_PX, orc.S_PY, orc.S_GX, orc.S_GY, orc.S_TH], i]
                assert (self.pools[mid[i]] == row[None, :]).all(axis=1).any(), tag + ' spawn tuple not from the pool'
        return int(done.sum()), int(o.is_crash.sum())


def _run(env, maps, n_pick, T, seed, sigma=0.02):
    B = env.B
    rng = np.random.RandomState(seed)
    pick = np.sort(rng.choice(B, n_pick, replace=False))
    sh = Shadow(env, maps, pick)
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    lo = torch.tensor([-0.05, -0.7], device='cuda')
    hi = torch.tensor([0.55, 0.7], device='cuda')
    n_done = n_crash = 0
    for t in range(T):
        act = lo + (hi - lo) * torch.rand(B, 2, device='cuda', generator=g)
        noise = sigma * torch.randn(B, 2, NB, device='cuda', generator=g)
        d, c = sh.step(act, noise, 'step %d' % t)
        n_done += d
        n_crash += c
    return n_done, n_crash


def _near(env, peds):
    """Move half of every environment's pedestrians to within 4 m of its robot, so that discs and
    segments shape the scans and crashes (-> re-scans, auto-resets) occur inside a short run."""
    B, P = peds.shape[:2]
    rows = env.state[:2].T.cpu().numpy()
    near = peds.copy()
    sel = np.random.RandomState(2).rand(B, P) < 0.5
    off = np.random.RandomState(3).uniform(-4, 4, (B, P, 2)).astype(np.float32)
    near[..., 0:2] = np.where(sel[..., None], rows[:, None, :].astype(np.float32) + off, peds[..., 0:2])
    near[..., 4:6] = near[..., 0:2]
    return near


def test_c3_parameters_match_oracle_on_a_subsample():
    """16 384 envs, 2000^2 outdoor map (max_range 4e6 cells), 20 pedestrians per env as legs +
    boxes (up to 40 discs / 80 segments), auto-reset: 96 shadowed environments, 24 steps; then
    the same batch without auto-reset, so that crash re-scans are compared too."""
    from nav_gym_b200 import worlds
    from nav_gym_b200.batched_env import BatchedNavGym
    B, P = 16384, 20
    m, mp, peds = worlds.c3_world('cuda:0', B, P, pool_n=8192)
    assert m['width'] == 2000 and m['height'] == 2000
    total = 0
    for auto in (True, False):
        env = BatchedNavGym(B, mp, device='cuda:0', seed=5, auto_reset=auto, record_hits=True)
        # start close to the pedestrians so that discs / segments shape the scans
        env.reset_from_spawn_pool(np.random.RandomState(1))
        env.attach_pedestrians(_near(env, peds))
        assert env.max_disc == 40 and env.max_seg == 80
        env.reset()
        n_done, n_crash = _run(env, [m], 96, 24 if auto else 12, seed=7 + auto)
        total += n_crash
        assert int(env._pnd.max()) > 20 and int(env._pns.max()) > 40
    assert total > 0   # crashes (pedestrians next to the robot) did occur in the sample


def test_c4_parameters_match_oracle_on_a_subsample():
    """8192 envs (one GPU's share of 65 536) over 8 indoor + 8 outdoor maps, 5..15 pedestrians,
    auto-reset that re-draws the map: 128 shadowed environments, 30 steps; half of the pedestrians
    start next to the robots, so that episodes end and restart on other maps inside the run."""
    from nav_gym_b200 import worlds
    from nav_gym_b200.batched_env import BatchedNavGym
    B = 8192
    ms, mp, map_id, peds, nped = worlds.c4_world('cuda:0', B, pool_n=2048)
    assert len(ms) == 16
    env = BatchedNavGym(B, mp, device='cuda:0', map_id=map_id, seed=6, auto_reset=True, resample_map=True,
                        record_hits=True)
    env.reset_from_spawn_pool(np.random.RandomState(2))
    env.attach_pedestrians(_near(env, peds), nped=nped)
    env.reset()
    mid0 = env.map_id.clone()
    n_done, n_crash = _run(env, ms, 128, 30, seed=9)
    assert n_done > 0
    assert int((env.map_id != mid0).sum()) > 0     # some environments moved to another map


def test_philox_noise_is_a_unit_normal_per_beam():
    """The production noise path (Philox4x32-10 + Box-Muller with fast intrinsics) as a
    distribution, 4096 envs x 512 beams on the bench world: beams at range_max get none
    (env.py:438-440); elsewhere (scan - clean scan) / sigma is N(0, 1): global mean / std,
    per-env mean and std, no correlation between neighbouring environments, between episodes
    (repeated resets advance the episode counter) or between steps (a robot standing still)."""
    from nav_gym_b200 import worlds
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool
    m, pool = worlds.load_bench_world()
    mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
    B, R = 4096, 6
    rng = np.random.RandomState(5)
    rows = pool[rng.randint(len(pool), size=B)]
    sigma = rng.uniform(0.01, 0.05, B).astype(np.float32)

    def make(sig):
        env = BatchedNavGym(B, mp, device='cuda:0', seed=77, auto_reset=False)
        env.set_state(rows[:, 0:2], rows[:, 2:4], rows[:, 4], noise_std=sig)
        return env
    clean = make(np.zeros(B, np.float32))
    c0 = clean.reset()[:, :NB].clone()
    noisy = make(sigma)
    sg = torch.from_numpy(sigma).cuda()[:, None]
    at_max = c0 == 25.0
    assert 0.005 < float(at_max.float().mean()) < 0.2
    zs, vs = [], []
    for r in range(R):                       # R episodes at the same pose
        s = noisy.reset()[:, :NB].clone()
        assert torch.equal(s[at_max], c0[at_max])            # no noise on beams at range_max
        zs.append(((s - c0) / sg).double())
        vs.append(~at_max)
    zero = torch.zeros(B, 2, device='cuda')
    for t in range(R):                       # R steps standing still (actions 0, 0)
        clean.step(zero)
        noisy.step(zero)                     # (a noisy rear beam may dip under its threshold: the
        c = clean.obs[:, :NB]                #  returned scan is then the re-scan, slot 1 -- as good)
        s = noisy.obs[:, :NB]
        assert not bool(clean.is_crash.any())
        assert torch.equal(s[c == 25.0], c[c == 25.0])
        zs.append(((s - c) / sg).double())
        vs.append(c != 25.0)
    Z, V = torch.stack(zs), torch.stack(vs)  # [2R, B, 512]
    Z = Z * V
    n = int(V.sum())
    zv = Z[V]
    assert abs(float(zv.mean())) < 5.0 / np.sqrt(n)
    assert abs(float(zv.std()) - 1.0) < 0.004                 # s.e. 1 / sqrt(2 n) ~ 1.5e-4
    assert abs(float((zv ** 3).mean())) < 0.02 and abs(float((zv ** 4).mean()) - 3.0) < 0.05
    assert float(zv.abs().max()) < 7.0
    # per environment
    cnt = V.sum((0, 2)).double()
    mean_e = Z.sum((0, 2)) / cnt
    std_e = torch.sqrt(((Z - mean_e[None, :, None]) ** 2 * V).sum((0, 2)) / (cnt - 1))
    assert float((mean_e.abs() * torch.sqrt(cnt)).max()) < 5.5
    assert float(((std_e - 1.0).abs() * torch.sqrt(2 * cnt)).max()) < 6.0
    # independence: neighbouring envs, consecutive episodes, consecutive steps (and episode vs step)
    def corr(a, b, mask):
        a, b = a[mask], b[mask]
        return float((a * b).mean() / (a.std() * b.std())), int(mask.sum())
    for k in range(Z.shape[0]):
        r_, n_ = corr(Z[k, :-1], Z[k, 1:], V[k, :-1] & V[k, 1:])
        assert abs(r_) < 5.0 / np.sqrt(n_), ('env / env+1', k, r_)
    for k in range(Z.shape[0] - 1):
        r_, n_ = corr(Z[k], Z[k + 1], V[k] & V[k + 1])
        assert abs(r_) < 5.0 / np.sqrt(n_), ('episode or step k / k+1', k, r_)
    r_, n_ = corr(Z[:, :, :-1], Z[:, :, 1:], V[:, :, :-1] & V[:, :, 1:])
    assert abs(r_) < 5.0 / np.sqrt(n_), ('beam / beam+1', r_)


def test_rollout_in_c_equals_device_stepping():
    """navgym_host_rollout (env groups rotated in C, pinned host buffers, action-bank policy)
    returns what device-resident stepping with the same actions returns: Philox noise and
    auto-reset are keyed by (seed, env, episode, step), not by the launch pattern."""
    from nav_gym_b200 import _lib, worlds
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool
    m, pool = worlds.load_bench_world()
    mp = MapPool([m], 'cuda:0', spawn_pools=[pool])
    B, K, rows_n = 1024, 37, 8
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    bank = torch.rand(rows_n, B, 2, device='cuda', generator=g) * torch.tensor([0.5, 1.28], device='cuda') \
        + torch.tensor([0, -0.64], device='cuda')
    envs = []
    for _ in range(2):
        env = BatchedNavGym(B, mp, device='cuda:0', seed=21, auto_reset=True, max_episode_steps=9)
        env.reset_from_spawn_pool(np.random.RandomState(8))
        envs.append(env)
    a, b = envs
    for s in range(K):
        a.step(bank[s % rows_n])
    torch.cuda.synchronize()
    act_h = bank.cpu().pin_memory()
    cur = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    obs_h = torch.empty(B, NB + 7, dtype=torch.float32).pin_memory()
    rew_h = torch.empty(B, dtype=torch.float32).pin_memory()
    done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    bounds = b.host_groups(3, cur, obs_h, rew_h, done_h)
    assert bounds[0][0] == 0 and bounds[-1][1] == B
    ab = _lib.ActionBank(C.c_void_p(act_h.data_ptr()), rows_n, B)
    b.rollout_host(K, C.cast(_lib.load().navgym_policy_action_bank, _lib.POLICY_FN), ab)
    assert torch.equal(obs_h, a.obs.cpu()) and torch.equal(rew_h, a.reward.cpu())
    assert torch.equal(done_h, a.done.cpu()) and torch.equal(b.state, a.state)
    assert torch.equal(b.episodes, a.episodes) and int(a.episodes.sum()) > B
    # a Python policy callable drives the same loop (GIL-bound convenience path)
    seen = []

    def policy(group, b0, b1, step):
        seen.append((group, step))
        cur[b0:b1] = act_h[step % rows_n, b0:b1]
    b.rollout_host(2, policy)
    a.step(bank[0])
    a.step(bank[1])
    torch.cuda.synchronize()
    assert torch.equal(obs_h, a.obs.cpu()) and len(seen) == 6


def test_non_square_map_and_origin_outside():
    """W != H: xy_to_ij clips i against height and j against width (env.py:1245-1248), so a robot
    beyond the short side keeps an origin cell outside the grid and every beam returns 'no hit'
    (range_libc's first bounds test); robots inside behave as usual.  Both orientations."""
    import cuda_util
    rng = np.random.RandomState(12)
    for H, W in ((120, 300), (300, 120)):
        data = np.zeros((H, W), np.int8)
        data[0, :] = data[-1, :] = 100
        data[:, 0] = data[:, -1] = 100
        data[H // 2 - 3:H // 2 + 3, W // 3:W // 3 + 8] = 100
        m = dict(data=data, origin=(0, 0), resolution=0.05, width=W, height=H)
        long_m = max(H, W) * 0.05
        start = np.array([[W * 0.025, H * 0.025], [1.0, 1.0], [W * 0.05 - 0.6, H * 0.05 - 0.6],
                          [long_m - 0.5, long_m - 0.5], [long_m + 2.0, 0.7], [0.7, long_m + 2.0]])
        B = len(start)
        goal = start + 2.0
        theta = rng.uniform(0, 2 * np.pi, B)
        map_id = np.zeros(B, np.int32)
        o = orc.OracleBatch([m], map_id, start, goal, theta, params=dict(t_stop=502.0))
        c = cuda_util.CudaStepper([m], map_id, start, goal, theta, early_stop=True)
        for s in (o, c):
            s.reset_obs()
        assert np.array_equal(c.obs[:, :NB], o.obs[:, :NB]), (H, W)
        for t in range(3):
            act = rng.uniform([0, -0.6], [0.5, 0.6], (B, 2)).astype(np.float32)
            for s in (o, c):
                s.step(act)
            assert np.array_equal(c.obs[:, :NB], o.obs[:, :NB]), (H, W, t)
            assert np.array_equal(c.hits, o.hits), (H, W, t)
            assert np.array_equal(c.done, o.done) and np.array_equal(c.is_crash, o.is_crash)
