"""Debug view of one environment (reference NavGymEnv.render, env.py:833-1058): the map with the
goal, the pedestrians and the robot with its footprints, and the lidar returns on top.

Host-side and off the hot path: it draws from the attribute surface of SURVEY 8b
(``map_info``, ``robot``, ``humans``, ``prev_obs``), which ``NavGymEnv`` keeps itself and
``BatchedNavGym.export_env(i)`` copies back from the device for one environment of a batch.
"""
import numpy as np

from .robot import KetiRobot, beam_table

SIZE = 800  # the reference shows an 800 x 800 window (env.py:834)


def _cell(xy, m):
    """World (x, y) -> integer (column, row) = (i, j) of the reference's xy_to_ij (env.py:1253)."""
    i = int(np.clip((xy[0] - m['origin'][0]) / m['resolution'], 0, m['width'] - 1))
    j = int(np.clip((xy[1] - m['origin'][1]) / m['resolution'], 0, m['height'] - 1))
    return i, j


def _polygon(img, cv2, m, px, py, theta, footprint, colour, thickness=1):
    c, s = np.cos(theta), np.sin(theta)
    pts = [_cell((c * x - s * y + px, s * x + c * y + py), m) for x, y in footprint]
    cv2.polylines(img, [np.array(pts, np.int32)], True, colour, thickness)


def draw(view, size=SIZE):
    """float32 [size, size, 3] image in [0, 1], row 0 = y min as in the reference (the map array
    is shown unflipped).  `view` needs map_info, robot, humans, prev_obs, num_scan_stack."""
    import cv2
    m, robot = view.map_info, view.robot
    res = float(m['resolution'])
    img = (np.asarray(m['data']) == 0).astype(np.float32)  # free white, occupied black
    img = cv2.merge([img, img, img])
    r = max(1, int(1.0 / res))
    i, j = _cell((robot.gx, robot.gy), m)
    cv2.rectangle(img, (i - r, j - r), (i + r, j + r), (0, 0, 1), -1)  # goal (env.py:843-851)
    for h in view.humans:
        rr = max(1, int(0.2 / res))
        i, j = _cell((h.gx, h.gy), m)
        cv2.rectangle(img, (i - rr, j - rr), (i + rr, j + rr), (1, 1, 0), -1)
    for h in view.humans:
        i, j = _cell((h.px, h.py), m)
        di, dj = int(0.6 * np.cos(h.theta) / res), int(0.6 * np.sin(h.theta) / res)
        cv2.arrowedLine(img, (i, j), (i + di, j + dj), (0, 0, 0), max(1, int(0.2 / res)))
        _polygon(img, cv2, m, h.px, h.py, h.theta, h.footprint, (0, 0, 0))
    i, j = _cell((robot.px, robot.py), m)
    di, dj = int(0.8 * np.cos(robot.theta) / res), int(0.8 * np.sin(robot.theta) / res)
    cv2.arrowedLine(img, (i, j), (i + di, j + dj), (1, 0, 0), max(1, int(0.2 / res)))
    _polygon(img, cv2, m, robot.px, robot.py, robot.theta, KetiRobot.footprint, (1, 0, 0))
    _polygon(img, cv2, m, robot.px, robot.py, robot.theta, KetiRobot.threshold_footprint, (0, 0, 1))
    _polygon(img, cv2, m, robot.px, robot.py, robot.theta, KetiRobot.discomfort_threshold_footprint, (0, 1, 0))
    if view.prev_obs is not None:  # newest scan of the stack, drawn where each beam ends
        n = KetiRobot.n_angles
        scan = np.asarray(view.prev_obs['observation'])[(view.num_scan_stack - 1) * n:view.num_scan_stack * n]
        ang = beam_table() + robot.theta
        for rng, a in zip(scan, ang):
            if 0 < rng < KetiRobot.range_max:
                cv2.circle(img, _cell((robot.px + rng * np.cos(a), robot.py + rng * np.sin(a)), m), 1, (1, 0, 1), -1)
    return cv2.resize(img, (size, size))


def render(view, mode='human'):
    """mode 'rgb_array' returns uint8 [800, 800, 3]; 'human' shows the window like the reference
    (needs a cv2 build with GUI support)."""
    img = draw(view)
    if mode == 'rgb_array':
        return (img * 255).astype(np.uint8)
    if mode == 'human':
        import cv2
        cv2.imshow('NavGym Env', img)
        cv2.waitKey(1)
        return None
    raise NotImplementedError(mode)
