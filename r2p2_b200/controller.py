"""Controller plugin base class -- same surface as the reference's (r2p2/controller.py:42-139).

The simulator side of the contract is unchanged: ``register_robot``, ``handle_collision``, ``set_dst``,
``set_ang``, ``control(dst) -> (speed, angular_velocity)``, ``on_collision``, ``write_info_to_log``,
``has_cur_detected_edge_list``, ``goal_oriented``.  For the two evolvable controllers the per-tick
``control`` runs fused inside the CUDA step kernel; the Python objects describe the controller (kind,
layer sizes, activation, genome) to the device and mirror its state back after every update.
"""
import math
from abc import ABC


class Controller(ABC):
    #: device controller kind understood by r2p2_set_controller, or None for host-only plugins
    device_kind = None

    def __init__(self, controller_type="", config=None):
        self.type = controller_type
        self.cur_detected_edges_distances = []
        self.cur_detected_edges = []
        self.actual_sensor_angles = []
        self.ang = []
        self.dst = []
        self.config = config
        self.robot = None

    def control(self, dst):
        """(speed, angular_velocity) for the distances ``dst``; the base class stands still."""
        self.update_sensor_angles(self.ang, dst)
        return 0, 0

    def register_robot(self, r):
        self.robot = r

    def write_info_to_log(self, log_file):
        pass

    def on_collision(self, pos):
        pass

    def handle_collision(self, col):
        pass

    def has_edge_list(self):
        return False

    def has_cur_detected_edge_list(self):
        return True

    def goal_oriented(self):
        return False

    def update_sensor_angles(self, angles, distances):
        self.cur_detected_edges_distances = distances
        self.actual_sensor_angles = angles
        for i in range(len(self.cur_detected_edges_distances)):
            if self.cur_detected_edges_distances[i] == math.inf:
                self.cur_detected_edges_distances[i] = 10000

    def set_dst(self, dst):
        self.dst = dst

    def set_ang(self, ang):
        self.ang = ang

    # ---- what the batched core needs to know (extensions; harmless for host-only plugins) ----
    def device_spec(self, n_rays):
        """dict(kind=..., hidden_layer=..., activation=..., genome=[...]) or None when the plugin has no device form."""
        return None


def upd_sensor_angles(function):
    """Decorator of the reference (controller.py:133-138): refresh the sensor angles before ``control``."""
    def wrapper(self, dst):
        self.update_sensor_angles(self.ang, dst)
        return function(self, dst)
    return wrapper
