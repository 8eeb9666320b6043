"""Sequential_PID_Controller -- the reference's goal-list PID controller (r2p2/controllers/pid_controller.py:8-262).

A host-side plugin: ``control(dst)`` runs between ``r2p2_sense`` and ``r2p2_move`` (group.RobotGroup, host mode), one
robot after the other as in the reference's loop.  Behaviour follows the reference statement by statement, quirks
included, because scenario files and downstream tools were tuned against them:

  * ``register_robot`` moves the robot to the first goal (pid_controller.py:255-259): the goal list starts with the
    start position;
  * the pair returned by ``control`` is unpacked by ``Robot.control`` as (speed, angular_velocity) (robot.py:385), so
    while backing off after a collision the robot is told ``speed = orientation`` and ``angular_velocity = -max_speed``
    (pid_controller.py:51-53);
  * ``choose_acceleration`` returns ``-max_acceleration`` for every value in (-max, max] (pid_controller.py:150-155);
  * in the avoid state a temporary goal is pushed in front of the list every tick (pid_controller.py:100-104);
  * ``is_at_goal`` reads the stage through the module-global ``utils.npdata`` (pid_controller.py:211) -- here the stage
    of the robot's group (``robot._group.env``), the same array.

Configuration: the reference reads a hard-coded ``../conf/controller-PID.json`` whatever dict it is given
(pid_controller.py:18); here the controller dict of the scenario is used (``config['controllers']``).
"""
import math

import numpy as np

from ..controller import Controller
from .neurocontroller import _controller_dict


class Sequential_PID_Controller(Controller):
    def __init__(self, config=None):
        super().__init__("SEQ_PID", config)
        f = _controller_dict(config, "Sequential_PID_Controller")
        if "goal" not in f and config is not None:
            f = next((c for c in config.get("controllers", []) if "goal" in c), f)
        self.goal = [list(g) if isinstance(g, (list, tuple)) else g for g in f["goal"]]
        self.ap, self.ai, self.ad = f["ap"], f["ai"], f["ad"]
        self.lp, self.li, self.ld = f["lp"], f["li"], f["ld"]
        self.accumulated_angle_error = 0
        self.accumulated_distance_error = 0
        self.last_angle_error = 0
        self.last_distance_error = 0
        self.max_acceleration = 5
        self.detected_edges = []
        self.cur_detected_edges = []
        self.actual_sensor_angles = []
        self.cur_detected_edges_distances = []
        self.target_angle = 0
        self.state = 0
        self.backing = False
        self.env = None                 # set explicitly, or taken from the robot's group (utils.npdata in the reference)

    # ---- driver (pid_controller.py:40-58) ------------------------------------------------------------
    def control(self, dst):
        self.update_sensor_angles(self.ang, dst)            # Controller.control (controller.py:62-68)
        if self.backing:
            self.backing -= 1
            return self.robot.orientation, -self.robot.max_speed
        self.manage_state(self.dst)
        if self.state == 0 or self.state == 2:
            return self.control_advance(self.ang, dst)
        return self.control_avoid(self.ang, dst)

    def control_advance(self, ang, dst):
        if self.is_done():
            return 0, 0
        angle = self.choose_angle(ang, dst)
        speed = self.robot.speed + self.choose_acceleration(ang, dst)
        if self.is_at_goal():
            self.switch_to_next_goal()
        return speed, angle

    def control_avoid(self, ang, dst):
        """pid_controller.py:73-124: push a temporary goal sideways of the obstacle, retreat if anything is very close."""
        rob = self.robot
        far = rob.vision_range[-1]
        corr_ang = 90
        risk = 1
        for d in dst:
            risk *= (d / far)
        if dst[1] + dst[2] < dst[-1] + dst[-2]:
            risk /= (far / dst[-1]) * (far / dst[-2])
        else:
            risk /= (far / dst[1]) * (far / dst[2])
        if dst[0] / far < 0.45:
            risk *= dst[0] / far
        safety_distance = (far + rob.radius) * risk * 2
        if safety_distance <= rob.vision_range[0]:
            safety_distance = rob.vision_range[0] * 2
        if (dst[1] + dst[2]) > (dst[-1] + dst[-2]):
            a = self.target_angle % 360 + corr_ang
        else:
            a = self.target_angle % 360 - corr_ang
        x_var = safety_distance * np.cos(np.radians(self.target_angle - a))
        y_var = safety_distance * np.sin(np.radians(self.target_angle - a))
        if a % 360 > 90 and a % 360 < 270:
            x_var *= -1
        if a % 360 > 180:
            y_var *= -1
        self.goal.insert(0, (rob.x + x_var, rob.y + y_var))
        self.calculate_angle_variation()                    # (updates the angular PID terms; its value is not used)
        self.target_angle = a
        must_retreat = False
        for d in self.dst:
            if d / far < rob.vision_range[0] / far * 12:
                must_retreat = True
                break
        speed = -rob.max_speed if must_retreat else rob.speed
        return speed, 0

    # ---- PID terms -----------------------------------------------------------------------------------
    def choose_angle(self, angles, distances):
        self.calculate_target_angle()
        return self.calculate_angle_variation()

    def choose_acceleration(self, angles, distances):
        acc = self.robot.acceleration + self.calculate_acceleration_variation()
        if acc > self.max_acceleration:
            return self.max_acceleration
        elif -acc < self.max_acceleration:
            return -self.max_acceleration
        return acc

    def calculate_acceleration_variation(self):
        if abs(self.target_angle - self.robot.orientation) > 45 or self.is_done():
            self.accumulated_distance_error += self.robot.brake()
            return self.robot.brake()
        e = np.linalg.norm((self.goal[0][0] - self.robot.x, self.goal[0][1] - self.robot.y))
        if self.robot.acceleration < self.max_acceleration:
            self.accumulated_distance_error += e
        de = e - self.last_distance_error
        self.last_distance_error = e
        return self.lp * e + self.li * self.accumulated_distance_error + self.ld * de

    def calculate_angle_variation(self):
        e = self.target_angle - self.robot.orientation
        self.accumulated_angle_error += e
        de = e - self.last_angle_error
        self.last_angle_error = e
        return self.ap * e + self.ai * self.accumulated_angle_error + self.ad * de

    def calculate_target_angle(self):
        self.target_angle = math.degrees(math.atan2(self.goal[0][1] - self.robot.y, self.goal[0][0] - self.robot.x))

    # ---- goals and state ---------------------------------------------------------------------------------
    def manage_state(self, dst):
        rng = self.robot.vision_range[-1]
        tolerance = self.robot.vision_range[0] / rng * 13
        if (dst[-1] / rng < tolerance and dst[-2] / rng < tolerance) or \
           (dst[1] / rng < tolerance and dst[2] / rng < tolerance) or dst[0] / rng < tolerance:
            self.state = 1 if self.state == 0 else 2
            return
        self.state = 0

    def _stage(self):
        if self.env is not None:
            return self.env
        g = getattr(self.robot, "_group", None)
        if g is None:
            raise RuntimeError("Sequential_PID_Controller: no stage yet (set controller.env or update the robot once)")
        return g.env

    def is_at_goal(self):
        if int(np.asarray(self._stage()).item((int(self.goal[0][0]), int(self.goal[0][1])))) == 0:
            return True
        return np.linalg.norm((self.goal[0][0] - self.robot.x, self.goal[0][1] - self.robot.y)) <= self.robot.radius * 1.85

    def is_done(self):
        return not self.goal

    def switch_to_next_goal(self):
        self.goal.pop(0)
        if self.state == 2:
            self.state = 0

    # ---- simulator callbacks ------------------------------------------------------------------------------
    def handle_collision(self, col):
        self.cur_detected_edges = col
        for e in col:
            if e not in self.detected_edges:
                self.detected_edges.append(e)

    def on_collision(self, pos):
        self.backing = 10

    def has_edge_list(self):
        return True

    def has_cur_detected_edge_list(self):
        return True

    def goal_oriented(self):
        return True

    def write_info_to_log(self, log_file):
        log_file.write("[GOAL: " + str(self.goal))
        log_file.write("; STATE: ADVANCING" if self.state == 0 else "; STATE: AVOIDING")
        log_file.write("; TARGET_ANGLE: " + str(self.target_angle) + "º]\n")

    def register_robot(self, r):
        self.robot = r
        first = self.goal[0]
        r.set_position(first[0], first[1])
        r.set_last_position(first[0], first[1])
