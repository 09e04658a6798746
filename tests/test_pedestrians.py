"""CPU: the pedestrian pipeline's pieces against what the reference computed (humans_*.npz,
minted by oracle/make_golden_humans.py from the unmodified env.py): the policy network, the
pedestrian lidar restated with the oracle, Human.set_vel, the footprint polygons."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import oracle as orc
from nav_gym_b200.robot import Human, beam_table


def oracle_human_scan(dist, m, pose, segs, cell_rule=1):
    """_compute_scan for a pedestrian (env.py:386-435) with the oracle's natives."""
    lx, ly, lt = np.float32(pose[0]), np.float32(pose[1]), np.float32(pose[2])
    lin = beam_table(Human)
    head, dirs = orc.beam_dirs(lt, lin)
    H, W = dist.shape
    ci = orc.lib().nvo_xy_to_cell(float(lx), float(m['origin'][0]), float(m['resolution']), H, cell_rule)
    cj = orc.lib().nvo_xy_to_cell(float(ly), float(m['origin'][1]), float(m['resolution']), W, cell_rule)
    ins = np.column_stack([np.full(512, ci), np.full(512, cj), head]).astype(np.float32)
    r = orc.calc_range_many(dist, ins, float(W * H)) * np.float32(m['resolution'])
    r = np.ascontiguousarray(r, np.float32)
    orc.render_segments(r, dirs, segs, np.array([lx, ly], np.float32))
    return np.clip(r, 0, np.float32(Human.range_max))


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_policy_network_reproduces_reference_means(name):
    """torch.manual_seed(1234); HumanPolicy() draws the weights the harness gave the reference;
    the same inputs give the same action means (same torch, CPU: bit-exact)."""
    from nav_gym_b200.pedestrians import HumanPolicy
    G = gu.load(name)
    torch.manual_seed(1234)
    pol = HumanPolicy()
    with torch.no_grad():
        for t in range(len(G['mean'])):
            x = torch.from_numpy(G['scan_in'][t])[:, None, :].expand(-1, 3, -1).contiguous()
            m = pol.mean(x, torch.from_numpy(G['goal_local'][t]), torch.from_numpy(G['speed'][t]))
            assert np.array_equal(m.numpy(), G['mean'][t])
            v, act, logp, m2 = pol(x, torch.from_numpy(G['goal_local'][t]), torch.from_numpy(G['speed'][t]))
            assert np.array_equal(m2.numpy(), G['mean'][t]) and v.shape == (x.shape[0], 1)


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_policy_inputs_follow_from_the_recorded_scans(name):
    """The policy sees the scan the pedestrian took at the end of the previous step, clipped to
    6 m and centred (env.py:627-629), and its own previous clipped mean as `speed`."""
    from nav_gym_b200.pedestrians import preprocess_scan
    G = gu.load(name)
    T = len(G['mean'])
    prev = G['scan0']
    for t in range(T):
        assert np.array_equal(preprocess_scan(torch.from_numpy(prev)).numpy(), G['scan_in'][t])
        prev = G['scan_out'][t]
        if t + 1 < T:
            assert np.array_equal(np.clip(G['mean'][t], [0, -1], [1, 1]).astype(np.float32), G['speed'][t + 1])
    assert not G['speed'][0].any()


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_oracle_pedestrian_scan_matches_reference(name):
    G = gu.load(name)
    m = gu.map_info(G)
    dist = orc.edt(np.asarray(m['data']) >= 0.1)
    for t in range(0, len(G['mean']), 7):
        for i in range(G['pose_after'].shape[1]):
            r = oracle_human_scan(dist, m, G['pose_after'][t, i], G['segs_out'][t, i, :G['nseg_out'][t, i]])
            assert np.array_equal(r, G['scan_out'][t, i]), (t, i)


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_set_vel_and_footprints(name):
    """Human.set_vel (human.py:32-41) restated in torch float64, and the closed footprints the
    other agents are drawn with (env.py:404-414) against the recorded segments."""
    from nav_gym_b200.pedestrians import footprint_polygons, human_set_vel
    G = gu.load(name)
    for t in range(len(G['mean'])):
        act = np.clip(G['mean'][t], [0, -1], [1, 1]).astype(np.float64) * G['v_pref'][:, None]
        pose, vel = human_set_vel(torch.from_numpy(G['pose_before'][t]), torch.from_numpy(act), 0.2)
        assert np.allclose(pose.numpy(), G['pose_after'][t], rtol=0, atol=1e-12)
        assert np.allclose(vel.numpy(), G['vel'][t], rtol=0, atol=1e-12)
        segs, skip = gu.human_env_segments(G, t)
        poly = footprint_polygons(torch.from_numpy(G['pose_after'][t]), Human.footprint).numpy()
        P = poly.shape[0]
        want = segs[5:].reshape(P, 5, 4)[:, :4]
        assert np.allclose(poly, want, rtol=0, atol=4e-6)
