"""Headless simulation driver: the reference's ``utils.load_simulation`` / ``create_robot(s)`` / ``update_loop``
(r2p2/utils.py:118-234, 273-280, 302-303) without pygame, batched on the device.

``Simulation`` groups the scenario's robots by (robot json, controller shape) -- ``group.RobotGroup`` -- and advances
every group with one kernel launch (device controllers) or one sense + one move launch (host controllers) per tick and
one read-back, instead of the reference's per-robot loop (utils.py:226-227, 233-234).
"""
import json

import numpy as np

from .config_manager import ConfigManager
from .controllers.controllers import get_controllers
from .group import RobotGroup, group_key
from .robot import Robot


def load_image(infilename):
    """PIL ``convert("L")`` -> int32 [H, W], exactly utils.load_image (r2p2/utils.py:273-280)."""
    from PIL import Image
    img = Image.open(infilename).convert("L")
    img.load()
    return np.asarray(img, dtype="int32")


def stage_env(npdata):
    """The transpose the reference applies before simulating (utils.py:302-303): env[x, y], shape (W, H)."""
    return np.ascontiguousarray(np.flipud(np.rot90(npdata)))


def create_robot(json_file, controller=None, device=0):
    """utils.create_robot (r2p2/utils.py:133-163): robot json -> Robot."""
    with open(json_file, "r") as fp:
        f = json.load(fp)
    return Robot(identifier=f["id"], x=f["x"], y=f["y"], orientation=f["orientation"],
                 vision_range=(f["sonar_range"][0], f["sonar_range"][1]), sensors=f["sonars"], radius=f["radius"],
                 max_speed=f["max_speed"], controller=controller, name=f.get("name"), device=device)


class Simulation:
    def __init__(self, config_mgr, device=0):
        self.cm = config_mgr if isinstance(config_mgr, ConfigManager) else ConfigManager(config_mgr)
        cfg = self.cm.get_config()
        self.config = cfg
        self.npdata = load_image(self.cm.resolve(cfg["stage"]))
        self.env = stage_env(self.npdata)
        ctrls = get_controllers(cfg)
        if not isinstance(ctrls, list):
            ctrls = [ctrls]
        paths = cfg["robot"] if isinstance(cfg["robot"], list) else [cfg["robot"]]
        self.robots = []
        for i, path in enumerate(paths):                       # utils.create_robots (utils.py:175-190)
            c = ctrls[i] if i < len(ctrls) else ctrls[0]
            self.robots.append(create_robot(self.cm.resolve(path), c, device=device))
        self.device = device
        self.groups = None

    def _make_groups(self):
        """Robots in scenario order, grouped by what a launch has to share; a group keeps that order."""
        buckets = {}
        for r in self.robots:
            buckets.setdefault(group_key(r), []).append(r)
        self.groups = [RobotGroup(rs, self.env, device=self.device) for rs in buckets.values()]

    def update(self, delta, write_stats=True):
        """One pass of ``for r in robots: r.update(npdata, delta, True)`` (utils.py:226-227, 233-234): one launch and one
        read-back per group."""
        if self.groups is None:
            self._make_groups()
        for g in self.groups:
            g.update(delta, write_stats)

    def run(self, ticks, delta=0.1):
        for _ in range(ticks):
            self.update(delta)
        return self.robots


def load_simulation(config_mgr, device=0):
    return Simulation(config_mgr, device=device)
