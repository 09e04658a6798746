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
