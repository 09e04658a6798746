"""CPU: the C oracle against the fixtures minted from the reference's own env.py."""
import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc


class OracleStepper(orc.OracleBatch):
    pass


def make_oracle(G, **params):
    md, ms = gu.geom_dims(G)
    p = dict(cell_rule=1)  # the fixtures were minted under NumPy 2 (float32 cell division)
    p['num_scan_stack'] = int(G['num_scan_stack']) if 'num_scan_stack' in G else 1
    p['min_turn_radius'] = float(G['min_turning_radius']) if 'min_turning_radius' in G else 0.0
    names = dict(time_step='dt', distance_threshold='dist_thresh', reward_scale='r_scale',
                 reward_success_factor='r_success', reward_crash_factor='r_crash',
                 reward_progress_factor='r_progress', reward_forward_factor='r_forward',
                 reward_rotation_factor='r_rotation', reward_discomfort_factor='r_discomfort')
    p.update({names[k]: v for k, v in gu.env_kwargs(G).items()})
    p.update(params)
    return OracleStepper([gu.map_info(G)], np.zeros(1, np.int32), G['start'][None, :2], G['goal'][None],
                         G['start'][2:3], params=p, max_disc=md, max_seg=ms)


@pytest.mark.parametrize('name', gu.trace_names())
def test_oracle_replays_reference_trace(name):
    G = gu.load(name)
    o = make_oracle(G)
    assert np.array_equal(o.thr, G['thr']) and np.array_equal(o.dthr, G['dthr'])
    T = gu.replay(G, o, pose_tol=1e-12, reward_tol=1e-12)
    assert T == len(G['actions'])


@pytest.mark.parametrize('name', gu.trace_names())
def test_early_stop_is_invisible_after_clip(name):
    """Stopping the march at 25 m / 0.05 + 2 = 502 cells changes no clipped range; hit cells
    beyond it become 'none'."""
    G = gu.load(name)
    o = make_oracle(G, t_stop=502.0)

    o.early_stop = True
    gu.replay(G, o, pose_tol=1e-12, reward_tol=1e-12)


def test_known_answers():
    K = dict(np.load(gu.GOLDEN + '/known_answers.npz'))
    # KA4 cell mapping, NumPy-2 rule as the reference computes it in the build container
    ij = np.array([[orc.lib().nvo_xy_to_cell(float(x), 0.0, 0.05, 1000, 1),
                    orc.lib().nvo_xy_to_cell(float(y), 0.0, 0.05, 1000, 1)] for x, y in K['cell_xy']])
    assert np.array_equal(ij, K['cell_ij'])
    # the NumPy-1 rule (float64 division) differs from it on a few cell-edge inputs only
    ij1 = np.array([[orc.lib().nvo_xy_to_cell(float(x), 0.0, 0.05, 1000, 0),
                     orc.lib().nvo_xy_to_cell(float(y), 0.0, 0.05, 1000, 0)] for x, y in K['cell_xy']])
    assert np.abs(ij1 - K['cell_ij']).max() <= 1
    assert (ij1 != K['cell_ij']).mean() < 0.1


def test_kinematics_known_answers():
    K = dict(np.load(gu.GOLDEN + '/known_answers.npz'))
    n = len(K['kin_state'])
    m = dict(data=np.zeros((1000, 1000), np.int8), origin=(0, 0), resolution=0.05, width=1000, height=1000)
    m['data'][0, :] = 100
    o = orc.OracleBatch([m], np.zeros(n, np.int32), K['kin_state'][:, :2], np.zeros((n, 2)),
                        K['kin_state'][:, 2])
    o.reset_obs()
    o.is_crash[:] = 0
    o.step(K['kin_action'].astype(np.float32))
    ok = o.is_crash == 0   # crashed envs are rolled back (env.py:707-723)
    assert ok.sum() > n // 2
    assert np.allclose(o.state[:3].T[ok], K['kin_out'][ok], rtol=0, atol=1e-12)


def test_third_edt_witness_scipy():
    """The oracle's Felzenszwalb-Huttenlocher EDT against scipy.ndimage.distance_transform_edt
    (an independent implementation): exact squared distances equal on reference-style maps."""
    from scipy import ndimage
    from nav_gym_b200 import maps
    rng = np.random.RandomState(5)
    for m in (maps.create_indoor_map(3, 60, rng, cells=60, scale=5), maps.create_outdoor_map(10, 0.7, rng, size=300),
              gu.map_info(gu.load(gu.trace_names()[0]))):
        occ = np.asarray(m['data']) >= 0.1
        d2 = orc.edt_sq(occ)
        want = ndimage.distance_transform_edt(~occ) ** 2
        assert np.array_equal(d2, np.rint(want).astype(np.int64))
        assert np.array_equal(orc.edt(occ), np.sqrt(np.rint(want)).astype(np.float32))


def test_axis_and_heading_convention_on_an_asymmetric_map():
    """What the code of the reference fixes about the native boundary (not its absent sources):
    map_info['data'] is [row = y][col = x] (map_generator.py:113-122; env.py:221,356 index
    data.T as [i = x, j = y]); the ray origin is (i, j) = (x cell, y cell) and beam k looks along
    lin[k] + theta in the world frame, counter-clockwise, beam 0 backwards and beam 256 forwards
    (env.py:386-390, 419-424; ros_env.py publishes the scan as a LaserScan from angle_min = -pi).
    One wall 3 m ahead on +x, another 5 m to the left on +y, nothing behind / to the right."""
    H, W = 400, 600
    data = np.zeros((H, W), np.int8)
    x0, y0 = 10.0, 8.0
    data[:, int((x0 + 3.0) / 0.05):int((x0 + 3.0) / 0.05) + 4] = 100          # wall at x = 13 m (columns)
    data[int((y0 + 5.0) / 0.05):int((y0 + 5.0) / 0.05) + 4, :] = 100          # wall at y = 13 m (rows)
    m = dict(data=data, origin=(0, 0), resolution=0.05, width=W, height=H)
    for theta, fwd, left in ((0.0, 3.0, 5.0), (np.pi / 2, 5.0, None)):
        o = orc.OracleBatch([m], np.zeros(1, np.int32), np.array([[x0, y0]]), np.array([[x0 + 1, y0]]),
                            np.array([theta]))
        scan = o.reset_obs()[0, :512]
        assert abs(scan[256] - fwd) < 0.06, (theta, scan[256])                # beam 256: straight ahead
        if left is not None:
            assert abs(scan[384] - left) < 0.06, scan[384]                    # beam 384: +90 deg = left
            assert scan[128] == 25.0 and scan[0] == 25.0                      # right / behind: open
        else:
            assert abs(scan[128] - 3.0) < 0.06                                # facing +y, the +x wall is on the right
            assert scan[384] == 25.0
    # a pedestrian box on the left shows up around beam 384, not 128 (render_contours_in_lidar)
    o = orc.OracleBatch([m], np.zeros(1, np.int32), np.array([[x0, y0]]), np.array([[x0 + 1, y0]]),
                        np.array([0.0]), max_seg=4)
    box = np.array([[[x0 - 0.2, y0 + 2.0, x0 + 0.2, y0 + 2.0], [x0 + 0.2, y0 + 2.0, x0 + 0.2, y0 + 2.4],
                     [x0 + 0.2, y0 + 2.4, x0 - 0.2, y0 + 2.4], [x0 - 0.2, y0 + 2.4, x0 - 0.2, y0 + 2.0]]], np.float32)
    scan = o.reset_obs(segs=box, nseg=np.array([4], np.int32))[0, :512]
    assert abs(scan[384] - 2.0) < 1e-3 and scan[128] == 25.0


def test_march_rounding_switch():
    """The oracle builds in two forms, x0 + dx * t fused (canonical) or separately rounded
    (-DNVO_MARCH_NO_FMA): the forms agree on almost every ray and are distinguishable, so that
    the real range_libc's choice can be matched by flipping the flag (the CUDA library has the
    same switch; tests/test_gpu_parity.py runs both pairs)."""
    import ctypes as C
    import os
    orc.build()
    alt = C.CDLL(os.path.join(os.path.dirname(orc.__file__), '_build', 'libnavgym_oracle_nofma.so'))
    assert orc.lib().nvo_march_is_fused() == 1 and alt.nvo_march_is_fused() == 0
    G = gu.load(gu.trace_names()[0])
    dist = orc.edt(np.asarray(G['map_data']) >= 0.1)
    H, W = dist.shape
    rng = np.random.RandomState(0)
    ys, xs = np.where(dist > 4)
    pick = rng.randint(len(xs), size=3000000)
    ins = np.column_stack([xs[pick], ys[pick], rng.uniform(-np.pi, np.pi, len(pick))]).astype(np.float32)
    a = orc.calc_range_many(dist, ins, float(W * H))
    b = np.empty(len(ins), np.float32)
    alt.nvo_calc_range_many(dist.ctypes.data_as(C.c_void_p), W, H, ins.ctypes.data_as(C.c_void_p),
                            b.ctypes.data_as(C.c_void_p), len(ins), C.c_float(W * H), C.c_float(W * H), None, None)
    diff = a != b
    # (the two roundings differ in 16 % of the sample positions but pick another CELL only ~1e-5 of
    # the time, and a ray's range changes more rarely still)
    assert 0 < diff.sum() and diff.mean() < 1e-3, diff.sum()


def test_native_call_transcript_fixture_replays_through_the_oracle():
    """tests/golden/native_calls.npz (the reference's own native-call sequence, minted by
    oracle/make_golden_native_calls.py) is self-consistent: the oracle's restatements reproduce
    every recorded result from the recorded arguments.  The GPU test replays the same transcript
    through the product's level-1 binding."""
    G = gu.load('native_calls')
    names = [str(x) for x in G['kind_names']]
    dist, max_range, seen = None, None, set()
    for i, kind in enumerate(G['kind']):
        fn = names[int(kind)]
        g = lambda k: G['c%d_%s' % (i, k)]
        seen.add(fn)
        if fn == 'PyRayMarching':
            H, W = [int(v) for v in g('occ_shape')]
            dist = orc.edt(np.unpackbits(g('occ_bits'))[:H * W].reshape(H, W).astype(np.bool_))
            max_range = float(g('max_range'))
        elif fn == 'calc_range_many':
            want, _ = orc.calc_range_many(dist, np.ascontiguousarray(g('ins')), max_range, want_hits=True)
            assert np.array_equal(want, g('outs'))
        else:
            r = g('ranges_in').copy()
            head = np.asarray(g('angles')).astype(np.float32)
            _, dirs = orc.beam_dirs(np.float32(0), lin=head.astype(np.float64))
            if fn == 'render_contours_in_lidar':
                orc.render_contours(r, dirs, g('flat'), np.asarray(g('lidar_xy'), np.float32))
            elif len(g('poses')):
                discs = np.concatenate([orc.legs_to_discs(p, s_) for p, s_ in zip(g('poses'), g('states'))]).astype(np.float32)
                orc.render_discs(r, dirs, discs, np.asarray(g('lidar_xy'), np.float32))
            assert np.array_equal(r, g('ranges_out'))
    assert seen == set(names)
