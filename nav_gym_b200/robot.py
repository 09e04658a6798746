"""Robot / pedestrian constants of NavGym-v0 (reference keti_robot.py:11-48, human.py:4-16).

Only the numbers the hot path needs; the integration itself (KetiRobot.set_vel,
keti_robot.py:64-93) runs inside the CUDA step kernel.
"""
import numpy as np


class KetiRobot(object):
    """Attribute surface that ros_env.py reads from ``env.robot`` (ros_env.py:87-151)."""
    footprint = [[0.3, 0.4], [-0.70, 0.4], [-0.70, -0.4], [0.3, -0.4]]
    threshold_footprint = [[0.6, 0.6], [-0.7, 0.6], [-0.7, -0.6], [0.6, -0.6]]
    discomfort_threshold_footprint = [[0.6 + 0.5, 0.6 + 0.5], [-0.7, 0.6 + 0.5],
                                      [-0.7, -0.6 - 0.5], [0.6 + 0.5, -0.6 - 0.5]]
    real_threshold_footprint = [[0.6, 0.6], [-1.0, 0.6], [-1.0, -0.6], [0.6, -0.6]]
    real_discomfort_threshold_footprint = [[0.6 + 1.0, 0.6 + 0.5], [-0.7, 0.6 + 0.5],
                                           [-0.7, -0.6 - 0.5], [0.6 + 1.0, -0.6 - 0.5]]
    has_legs = False
    angle_increment = 0.0122718463
    angle_min = -3.141592
    angle_max = 3.141592
    range_max = 25.
    n_angles = 512
    rotation_centre_offset = 0.14474  # keti_robot.py:72-73

    def __init__(self, px, py, theta, gx, gy, time_step):
        self.px, self.py, self.theta = px, py, theta
        self.gx, self.gy = gx, gy
        self.time_step = time_step
        self.vx, self.vy, self.v, self.r = 0., 0., 0., 0.


class Human(object):
    footprint = [[0.22, 0.19], [-0.22, 0.19], [-0.22, -0.19], [0.22, -0.19]]
    has_legs = True
    angle_increment = 0.00613592315
    angle_min = -1.57079632679
    angle_max = 1.57079632679
    range_max = 6.
    n_angles = 512

    def __init__(self, px, py, theta, gx, gy, time_step):
        self.px, self.py, self.theta = px, py, theta
        self.gx, self.gy = gx, gy
        self.time_step = time_step
        self.vx, self.vy, self.v, self.r = 0., 0., 0., 0.


def beam_table(agent=KetiRobot):
    """Beam angles before the heading is added (env.py:388-390), float64[n_angles]."""
    return np.linspace(agent.angle_min, agent.angle_max - agent.angle_increment, agent.n_angles)


def closed_segments(polygon):
    """[V,2] polygon -> [V,4] segments with the closing edge (pymap2d flatten_contours links
    the last vertex back to the first)."""
    p = np.asarray(polygon, np.float32).reshape(-1, 2)
    return np.concatenate([p, np.roll(p, -1, axis=0)], axis=1).astype(np.float32)


def footprint_segments(px, py, theta, footprint):
    """World-frame closed footprint of an agent as segments (env.py:408-414, utils.py:48-63)."""
    fp = np.asarray(footprint, np.float64)
    c, s = np.cos(theta), np.sin(theta)
    w = np.column_stack([c * fp[:, 0] - s * fp[:, 1] + px, s * fp[:, 0] + c * fp[:, 1] + py])
    return closed_segments(w)


def legs_to_discs(pose, dist_travelled):
    """pymap2d CSimAgent 'legs' (call site env.py:398-402): two leg discs of radius 0.03 m whose
    fore-aft / lateral offsets swing with the distance travelled in the base frame."""
    x, y, th = [float(v) for v in pose]
    s = [float(v) for v in dist_travelled]
    leg_radius, side_off, side_amp, front_amp = 0.03, 0.1, 0.1, 0.3
    front = front_amp * np.cos(s[0] * 2.0 / front_amp + s[2])
    side = side_amp * np.cos(s[1] * 2.0 / side_amp + s[2])
    out = np.empty((2, 3), np.float64)
    for i, (lx, ly) in enumerate(((front, side + side_off), (-front, -side - side_off))):
        out[i, 0] = x + np.cos(th) * lx - np.sin(th) * ly
        out[i, 1] = y + np.sin(th) * lx + np.cos(th) * ly
        out[i, 2] = leg_radius
    return out.astype(np.float32)
