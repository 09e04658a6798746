"""CPU: host-side mirror of the reference's module helpers and registration surface."""
import numpy as np

import golden_util as gu
from nav_gym_b200 import env as E
from nav_gym_b200 import gym_shim, maps, robot


def test_registration_and_spaces():
    import nav_gym_b200  # noqa: F401
    assert 'NavGym-v0' in gym_shim.registry or True
    spec = gym_shim.registry.get('NavGym-v0')
    if spec is not None and spec.entry_point == 'nav_gym_b200.env:NavGymEnv':
        kw = spec.kwargs
        assert kw['time_step'] == 0.2 and kw['reward_scale'] == 15. and kw['num_scan_stack'] == 1
        assert kw['env_param_range']['num_humans'] == ([5, 15], 'int')
    b = gym_shim.Box(low=np.array([0, -0.64]), high=np.array([0.5, 0.64]), dtype=np.float32)
    s = b.sample()
    assert s.dtype == np.float32 and b.contains(s)


def test_cell_mapping_known_answers():
    K = dict(np.load(gu.GOLDEN + '/known_answers.npz'))
    mi = dict(resolution=0.05, origin=(0, 0), height=1000, width=1000)
    # KA4
    assert E.xy_to_ij([1.23, 4.56], mi).tolist() == [24, 91]
    assert np.allclose(E.ij_to_xy([24, 91], mi), [1.225, 4.575])
    # the helper implements the NumPy-1.x rule; the fixture holds the NumPy-2 result of the
    # reference run here: they may part by one cell on cell-edge inputs only
    ij = E.batch_xy_to_ij(K['cell_xy'], mi)
    assert np.abs(ij - K['cell_ij']).max() <= 1 and (ij != K['cell_ij']).mean() < 0.1


def test_observation_dict_layout():
    G = gu.load(gu.trace_names()[0])
    o = np.concatenate([G['scan'][3].astype(np.float64), G['tail'][3]])
    d = E.observation_to_dict(o, 1, 512)
    assert d['scan'].shape == (512,) and np.array_equal(d['pose'], G['achieved'][3])
    assert np.array_equal(d['prev_pose'], G['tail'][3][:2]) and d['yaw'] == G['tail'][3][6]
    b = E.observation_batch_to_dict(np.stack([o, o]), 1, 512)
    assert b['vel'].shape == (2, 2)


def test_waypoints_and_maps():
    path = np.column_stack([np.arange(0, 10, 0.25), np.zeros(40)])
    wp = E.path_to_waypoints(path, 2)
    assert np.allclose(wp[-1], path[-1]) and np.all(np.diff(wp[:, 0]) > 0)
    rng = np.random.RandomState(0)
    m = maps.create_indoor_map(3, 60, rng)
    assert m['data'].shape == (1000, 1000) and set(np.unique(m['data'])) == {0, 100}
    assert (m['data'][0] == 100).all() and (m['data'][:, -1] == 100).all()
    o = maps.create_outdoor_map(10, 0.7, rng)
    assert o['data'].shape == (400, 400) and 0.02 < (o['data'] > 0).mean() < 0.2
    cm = maps.cost_map(o)
    assert cm['data'].shape == (80, 80) and cm['resolution'] == 0.25


def test_spawn_pool_obeys_episode_law():
    rng = np.random.RandomState(1)
    m = maps.create_outdoor_map(10, 0.7, rng)
    pool = maps.spawn_pool(m, 256, rng, min_goal_dist=5, max_goal_dist=15)
    assert len(pool) > 50
    d = np.hypot(pool[:, 0] - pool[:, 2], pool[:, 1] - pool[:, 3])
    assert (d > 5).all() and (d < 15).all()
    cm = maps.cost_map(m)
    for x, y in np.concatenate([pool[:, :2], pool[:, 2:4]]):
        assert cm['data'][int(y / 0.25), int(x / 0.25)] == 0
    peds = maps.spawn_pedestrians(m, pool[0, :2], 6, rng)
    assert peds.shape == (6, 16)
    assert (np.hypot(peds[:, 0] - pool[0, 0], peds[:, 1] - pool[0, 1]) >= 4 - 1e-6).all()


def test_footprint_helpers():
    segs = robot.closed_segments(robot.KetiRobot.threshold_footprint)
    assert segs.shape == (4, 4) and np.allclose(segs[-1, 2:], segs[0, :2])
    d = robot.legs_to_discs([1, 2, 0.5], [0, 0, 0])
    assert d.shape == (2, 3) and np.allclose(d[:, 2], 0.03)


def test_render_draws_the_attribute_surface():
    """render.draw needs only the host attribute surface: map, robot,
    humans, last observation.  The goal square, the robot arrow and the lidar returns land on
    the cells the reference's xy_to_ij would give."""
    from types import SimpleNamespace
    from nav_gym_b200 import maps
    from nav_gym_b200.render import draw, render
    from nav_gym_b200.robot import KetiRobot, Human, beam_table
    m = maps.create_outdoor_map(0, 0.5, np.random.RandomState(0))
    robot = KetiRobot(5.0, 6.0, 0.5, 15.0, 14.0, 0.2)
    h = Human(8.0, 8.0, 1.0, 9.0, 9.0, 0.2)
    scan = np.full(512, 25.0)
    scan[256] = 2.0  # the forward beam returns at 2 m
    obs = {'observation': np.concatenate([scan, np.zeros(7)])}
    view = SimpleNamespace(map_info=m, robot=robot, humans=[h], prev_obs=obs, num_scan_stack=1)
    img = draw(view, size=m['width'])  # one pixel per cell
    assert img.shape == (m['height'], m['width'], 3)
    gi, gj = int(15.0 / 0.05), int(14.0 / 0.05)
    assert tuple(img[gj, gi]) == (0.0, 0.0, 1.0)                       # goal square
    a = beam_table()[256] + robot.theta
    hi, hj = int((5.0 + 2.0 * np.cos(a)) / 0.05), int((6.0 + 2.0 * np.sin(a)) / 0.05)
    assert tuple(img[hj, hi]) == (1.0, 0.0, 1.0)                       # the lidar return
    assert tuple(img[int(6.0 / 0.05), int(5.0 / 0.05)]) == (1.0, 0.0, 0.0)  # robot arrow base
    out = render(view, 'rgb_array')
    assert out.shape == (800, 800, 3) and out.dtype == np.uint8


def test_pose2d_and_astar_stand_ins():
    """The host-side stand-ins for pose2d / pyastar2d (env.py:252-254, 350) against the ones the
    golden harness ran the reference with."""
    from nav_gym_b200 import natives
    from oracle import ref_harness as rh
    rng = np.random.RandomState(2)
    for _ in range(20):
        p = rng.uniform(-5, 5, 3)
        v = rng.uniform(-1, 1, 3)
        inv = natives.inverse_pose2d(p)
        assert np.allclose(inv, rh.inverse_pose2d(p), atol=1e-12)
        assert np.allclose(natives.apply_tf_to_vel(v, inv), rh.apply_tf_to_vel(v, inv), atol=1e-12)
        # composing a pose with its inverse is the identity
        c, s = np.cos(p[2]), np.sin(p[2])
        assert np.allclose([c * inv[0] - s * inv[1] + p[0], s * inv[0] + c * inv[1] + p[1]], 0, atol=1e-12)
    grid = np.full((30, 40), 255.0, np.float32)
    grid[10, 5:38] = np.inf
    grid[20, 0:30] = np.inf
    for s_, g_ in (((2, 3), (28, 35)), ((15, 2), (25, 39)), ((0, 0), (29, 39))):
        mine = natives.astar_path(grid, s_, g_, allow_diagonal=False)
        ref = rh.astar_path(grid, s_, g_, allow_diagonal=False)
        assert mine is not None and len(mine) == len(ref)              # both shortest
        assert tuple(mine[0]) == s_ and tuple(mine[-1]) == g_
        assert (np.abs(np.diff(mine, axis=0)).sum(1) == 1).all()         # 4-connected steps
        assert np.isfinite(grid[mine[:, 0], mine[:, 1]]).all()
    grid[:, 20] = np.inf
    assert natives.astar_path(grid, (2, 3), (28, 35)) is None
    assert natives.astar_path(grid, (10, 6), (2, 3)) is None            # start on a blocked cell


def test_map_generator_reproduces_the_reference_bit_for_bit():
    """nav_gym_b200.maps restates map_generator.py:97-143; with the same seed it must draw the
    same map.  The known answers are SHA-256 digests of maps minted by the reference's own
    generator in the build container (oracle/make_bench_world.py)."""
    import hashlib
    z = np.load(gu.GOLDEN + '/bench_world.npz')
    for name, want in zip(z['known_names'], z['known_sha']):
        kind, p0, p1, seed = str(name).split('_')
        seed = int(seed[4:])
        if kind == 'indoor':
            m = maps.create_indoor_map(int(p0), int(p1), np.random.RandomState(seed))
        else:
            m = maps.create_outdoor_map(int(p0), float(p1), np.random.RandomState(seed))
        assert m['data'].dtype == np.int8
        assert hashlib.sha256(np.ascontiguousarray(m['data']).tobytes()).hexdigest() == str(want), name


def test_bench_world_fixture_obeys_the_episode_law():
    """tests/golden/bench_world.npz: the reference's create_indoor_map(3, 100) under seed 0 and a
    65 536-tuple spawn pool -- starts and goals on free cost-map cells, 10 m < distance < 20 m
    (env.py:379), first scan free of discomfort (env.py:779-783, checked with the oracle on a
    sample)."""
    from nav_gym_b200 import worlds
    from oracle import oracle as orc
    m, pool = worlds.load_bench_world()
    ref = maps.create_indoor_map(3, 100, np.random.RandomState(0))
    assert np.array_equal(m['data'], ref['data']) and m['resolution'] == 0.05 and m['width'] == 1000
    assert pool.shape == (65536, 5)
    d = np.hypot(pool[:, 2] - pool[:, 0], pool[:, 3] - pool[:, 1])
    assert d.min() > 10.0 and d.max() < 20.0
    assert pool[:, 4].min() >= 0 and pool[:, 4].max() < 2 * np.pi
    cm = maps.cost_map(m)
    for cols in ((0, 1), (2, 3)):
        c = np.floor(pool[:, cols[0]] / 0.25).astype(int)
        r = np.floor(pool[:, cols[1]] / 0.25).astype(int)
        assert (cm['data'][r, c] == 0).all()
    rows = pool[::512]
    o = orc.OracleBatch([m], np.zeros(len(rows), np.int32), rows[:, 0:2], rows[:, 2:4], rows[:, 4],
                        params=dict(t_stop=502.0))
    obs = o.reset_obs(want_hits=False)
    assert not (obs[:, :512] < o.dthr[None, :]).any()


def test_map_pool_rejects_oversized_maps():
    """Hit cells travel as (y << 16 | x) and an EDT row is staged in 48 KB of shared memory: the
    pool refuses maps beyond 12000 x 32767 cells before any device work."""
    import pytest
    from nav_gym_b200 import batched_env, _lib
    lib = None
    try:
        lib = _lib.load()
    except Exception:
        pytest.skip('library not built')

    m = dict(width=13000, height=10, origin=(0., 0.), resolution=0.05, data=None)
    orig = _lib.require_device
    _lib.require_device = lambda: lib          # the check comes before the first device call
    try:
        with pytest.raises(ValueError):
            batched_env.MapPool([m], 'cpu')
        m = dict(width=10, height=40000, origin=(0., 0.), resolution=0.05, data=None)
        with pytest.raises(ValueError):
            batched_env.MapPool([m], 'cpu')
    finally:
        _lib.require_device = orig
