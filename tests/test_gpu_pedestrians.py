"""GPU: the pedestrian pipeline on the device against the reference's recorded pipeline
(humans_*.npz) -- the batched pedestrian lidar bit-exactly, the policy forward within float32
summation-order tolerance."""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


def _scanner(G, B, P, max_seg):
    from nav_gym_b200.batched_env import MapPool
    from nav_gym_b200.pedestrians import AgentScanner
    pool = MapPool([gu.map_info(G)], 'cuda:0')
    map_id = torch.zeros(B, dtype=torch.int32, device='cuda:0')
    return AgentScanner(pool, B, P, max_seg, map_id, cell_rule='numpy2')


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_agent_scan_batch_matches_reference_scans(name):
    """Every recorded step becomes one environment of a batch: the pedestrians' scans from their
    recorded poses, against the map and the recorded footprints of the others, equal the scans
    the reference stored -- bit for bit."""
    G = gu.load(name)
    T, P = G['pose_after'].shape[:2]
    per = [gu.human_env_segments(G, t) for t in range(T)]
    S = per[0][0].shape[0]
    segs = torch.from_numpy(np.stack([p[0] for p in per])).cuda()
    skip = torch.from_numpy(np.stack([p[1] for p in per])).cuda()
    nseg = torch.full((T,), S, dtype=torch.int32, device='cuda')
    sc = _scanner(G, T, P, S)
    r = sc.scan(torch.from_numpy(G['pose_after']).cuda().contiguous(), segs, nseg, skip)
    torch.cuda.synchronize()
    assert np.array_equal(r.cpu().numpy(), G['scan_out'])
    # fewer live agents than slots: the dead slots' rows are left alone
    sc.ranges.fill_(-1.0)
    nag = torch.full((T,), P - 1, dtype=torch.int32, device='cuda')
    r = sc.scan(torch.from_numpy(G['pose_after']).cuda().contiguous(), segs, nseg, skip, nag)
    torch.cuda.synchronize()
    r = r.cpu().numpy()
    assert np.array_equal(r[:, :P - 1], G['scan_out'][:, :P - 1]) and (r[:, P - 1] == -1).all()
    # the first scans of the episode (reset, env.py:808-815) from the start poses
    segs0, skip0 = gu.human_env_segments(G, 0)
    from nav_gym_b200.pedestrians import footprint_polygons
    from nav_gym_b200.robot import Human, KetiRobot
    poly = footprint_polygons(torch.from_numpy(G['pose0']), Human.footprint).reshape(-1, 4)
    rob = footprint_polygons(torch.from_numpy(G['robot0']), KetiRobot.threshold_footprint).reshape(-1, 4)
    env_segs = torch.cat((rob, poly)).cuda()[None].contiguous()
    sk = torch.tensor([[4 * (1 + j), 4] for j in range(P)], dtype=torch.int32, device='cuda')[None].contiguous()
    sc1 = _scanner(G, 1, P, env_segs.shape[1])
    r0 = sc1.scan(torch.from_numpy(G['pose0'])[None].cuda().contiguous(), env_segs,
                  torch.tensor([env_segs.shape[1]], dtype=torch.int32, device='cuda'), sk)
    torch.cuda.synchronize()
    got, want = r0[0].cpu().numpy(), G['scan0']
    # footprints rebuilt on the device may differ from the reference's by a float32 ulp, which
    # moves a beam's hit on a footprint edge by ~1e-6 m; map hits stay exact
    assert np.allclose(got, want, rtol=0, atol=2e-5) and (got == want).mean() > 0.97


@pytest.mark.parametrize('name', gu.human_trace_names())
def test_policy_forward_on_device(name):
    from nav_gym_b200.pedestrians import HumanPolicy
    G = gu.load(name)
    torch.manual_seed(1234)
    pol = HumanPolicy().cuda()
    T, P = G['mean'].shape[:2]
    x = torch.from_numpy(G['scan_in']).cuda().reshape(T * P, 1, 512).expand(-1, 3, -1).contiguous()
    with torch.no_grad():
        m = pol.mean(x, torch.from_numpy(G['goal_local']).cuda().reshape(-1, 2),
                     torch.from_numpy(G['speed']).cuda().reshape(-1, 2))
    assert np.allclose(m.cpu().numpy().reshape(T, P, 2), G['mean'], rtol=0, atol=2e-5)
