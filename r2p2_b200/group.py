"""RobotGroup -- ``for r in robots: r.update(npdata, delta, True)`` (r2p2/utils.py:226-227, 233-234) as ONE launch.

The robots of a scenario that share the stage, the robot json and the controller shape form a group: one ``BatchSim``
holds their state on the device, a tick is one ``r2p2_run(h, 1, ...)`` (controllers with a device form: the MLP / linear
controllers run fused in the step kernel) or one ``r2p2_sense`` + one ``r2p2_move`` (any other ``Controller`` subclass:
``control(dst)`` runs on the host in between, robot by robot, exactly where the reference calls it, robot.py:385), and
one read-back of state and trace for the whole group.  The ``Robot`` objects stay what callers and controllers see:
pose attributes are pushed to the device when host code assigned them (``set_position``, ``robot.orientation = ...``:
``r2p2_write_pose``, odometry untouched) and refreshed after every tick; ``Robot.collide`` / ``controller.on_collision``
are called with the spots the reference passes (robot.py:211-253), in its order.

Sensor noise (``Robot.has_noise``, the reference's default): the reference draws ``np.random.normal(0, 0.5)`` once per
unclamped ray, robot after robot, from numpy's global stream.  Which rays are unclamped does not depend on the noise, so
a noise-free sensor sweep fixes every robot's draw count, the global stream is cut accordingly, and the device replays
each robot's slice: a seeded run of a robot list reproduces the reference draw for draw.
"""
import math

import numpy as np

from .logs import collision_spots, format_collision, python_dst
from .sim import BatchSim


def ray_angles(sensors):
    """Ray angles with the reference's own Python types (JSON list entries / range() ints, utils.py:465-466)."""
    if isinstance(sensors, (list, tuple)):
        return list(sensors)
    return list(range(0, 360, math.floor(365 / int(sensors))))


def group_key(robot):
    """Robots with equal keys can share a launch: same ray fan, ranges, radius, speed limit and controller shape."""
    sensors = tuple(robot.sensors) if isinstance(robot.sensors, (list, tuple)) else int(robot.sensors)
    n_rays = len(ray_angles(robot.sensors))
    spec = robot.controller.device_spec(n_rays)
    if spec is None:
        shape = ("host",)
    else:
        c = robot.controller
        shape = (spec["kind"], tuple(spec["hidden_layer"] or ()), spec["activation"], bool(getattr(c, "evolve", False)),
                 float(getattr(c, "time", 0) or 0))
    return (sensors, tuple(robot.vision_range), robot.radius, robot.max_speed, shape)


class RobotGroup:
    def __init__(self, robots, env, device=0):
        self.robots = list(robots)
        assert self.robots and len({group_key(r) for r in self.robots}) == 1, "robots of a group must share robot json and controller shape"
        r0 = self.robots[0]
        self.env = env
        self.sim = BatchSim(device)
        self.sim.set_stage(np.asarray(env))
        self.sim.set_robot(sonar_range=r0.vision_range, radius=r0.radius, max_speed=r0.max_speed, sonars=r0.sensors)
        self.R = self.sim.R
        self.host_mode = r0.controller.device_spec(self.R) is None
        self.clamp = r0.vision_range[1] + r0.radius
        self._started = False
        self._mirror = [None] * len(self.robots)          # (x, y, orientation, last_pos) as the device holds them
        for r in self.robots:
            r._group = self

    def close(self):
        self.sim.close()

    # ---- host <-> device ---------------------------------------------------------------------------
    def _genomes(self):
        if self.host_mode:
            return np.zeros((len(self.robots), 2))
        return np.asarray([r.controller.device_spec(self.R)["genome"] for r in self.robots], dtype=np.float64)

    def _start(self):
        sim, r0 = self.sim, self.robots[0]
        if self.host_mode:
            sim.set_controller("const")
        else:
            spec = r0.controller.device_spec(self.R)
            sim.set_controller(spec["kind"], hidden_layer=spec["hidden_layer"], activation=spec["activation"])
        sim.upload_genomes(self._genomes())
        budget = float(getattr(r0.controller, "time", 0) or 0)
        sim.reset([float(r.x) for r in self.robots], [float(r.y) for r in self.robots], [float(r.orientation) for r in self.robots],
                  time_budget=budget)
        lx, ly = [float(r.last_pos[0]) for r in self.robots], [float(r.last_pos[1]) for r in self.robots]
        sim.write_pose(0, last_x=lx, last_y=ly)
        sim.set_trace(True, 1)
        for i, r in enumerate(self.robots):
            self._mirror[i] = (r.x, r.y, r.orientation, tuple(r.last_pos))
            r._genome_dirty = False
        self._started = True

    def _push(self):
        """Poses / last positions / genomes host code changed since the last tick."""
        if not self._started:
            self._start()
            return
        sim = self.sim
        if not self.host_mode and any(getattr(r, "_genome_dirty", False) for r in self.robots):
            sim.upload_genomes(self._genomes())
            for r in self.robots:
                r._genome_dirty = False
        for i, r in enumerate(self.robots):
            cur = (r.x, r.y, r.orientation, tuple(r.last_pos))
            m = self._mirror[i]
            if cur != m:
                sim.write_pose(i, x=[float(r.x)], y=[float(r.y)], theta=[float(r.orientation)],
                               last_x=[float(r.last_pos[0])], last_y=[float(r.last_pos[1])])
                self._mirror[i] = cur

    def _noise_streams(self, clean_dst):
        """Cut numpy's global legacy stream the way the reference consumes it: robot after robot, one draw per unclamped
        ray.  Returns [P, R] per-robot draws (zeros where a robot has no noise or the ray is clamped)."""
        P, R = clean_dst.shape
        streams = np.zeros((P, R))
        cf = float(self.clamp)
        for i, r in enumerate(self.robots):
            if not r.has_noise:
                continue
            n = int((clean_dst[i] != cf).sum())
            if n:
                streams[i, :n] = np.random.normal(0, 0.5, size=n)
        return streams

    def _pull(self, theta_before, last_before, write_stats):
        sim = self.sim
        s = sim.state()
        tr = sim.trace(1)
        tr_dst = tr["dst"][0]
        fit = None if self.host_mode else sim.fitness()
        for i, r in enumerate(self.robots):
            r.x, r.y, r.orientation = float(s["x"][i]), float(s["y"][i]), float(s["theta"][i])
            r.speed, r.angular_velocity = float(s["speed"][i]), float(s["omega"][i])
            r.flags = int(s["flags"][i])
            ctrl = r.controller
            if not self.host_mode:
                dst = [float(d) for d in tr_dst[i]]
                ctrl.set_dst(dst)
                ctrl.set_ang([a + theta_before[i] for a in ray_angles(r.sensors)])
                ctrl.update_sensor_angles(ctrl.ang, list(dst))
                ctrl._fitness = float(fit[i])
                if r._logs is not None:
                    r._logs.sensor_data(r.identifier, ctrl.ang, theta_before[i], python_dst(dst, self.clamp))
            cinfo = tr["cinfo"][0, i]
            spots = collision_spots(cinfo, (sim.W, sim.H), last_before[i])
            for spot in spots:                                # Robot.collide (robot.py:170-179): log line + controller.on_collision
                if r._logs is not None:
                    r._logs.log.write(format_collision(r.identifier, spot))
                r.collide(spot)
            if not (int(cinfo[0]) & 32):                      # robot.py:249-254: put back to last_pos, or last_pos = (x, y)
                r.last_pos = (r.x, r.y)
            self._mirror[i] = (r.x, r.y, r.orientation, tuple(r.last_pos))
            if write_stats and r._logs is not None:
                r._logs.stats(r)
        bad = [r.identifier for r in self.robots if r.flags & 2]
        if bad:
            raise IndexError("robot(s) %s left the map or have a non-finite pose (the reference raises here too)" % bad)

    # ---- one tick for every robot of the group --------------------------------------------------------
    def update(self, delta, write_stats=False):
        self._push()
        sim = self.sim
        theta_before = [r.orientation for r in self.robots]
        last_before = [tuple(r.last_pos) for r in self.robots]
        noisy = any(r.has_noise for r in self.robots)
        if self.host_mode:
            sim.set_noise("off")
            dst_all, px = sim.sense(hit_pixels=True)          # Robot.check_sensors for everybody
            if noisy:
                dst_all = dst_all.copy()
                cf = float(self.clamp)
                for i, r in enumerate(self.robots):           # robot.py:308-310, in robot order
                    if r.has_noise:
                        for k in range(self.R):
                            if dst_all[i, k] != cf:
                                dst_all[i, k] += np.random.normal(0, 0.5)
            cmds = np.zeros((len(self.robots), 2))
            for i, r in enumerate(self.robots):
                ctrl = r.controller
                dst = python_dst(dst_all[i], self.clamp)
                ang = [a + theta_before[i] for a in ray_angles(r.sensors)]
                col = [(float(p[0]), float(p[1])) for p in px[i] if p[0] >= 0]
                if r._logs is not None:
                    r._logs.sensor_data(r.identifier, ang, theta_before[i], dst)
                ctrl.handle_collision(col)                    # robot.py:320-323, 379
                ctrl.set_dst(dst)
                ctrl.set_ang(ang)
                speed, angular_velocity = ctrl.control(dst)   # robot.py:385
                cmds[i] = (float(speed), float(angular_velocity))
            sim.move(cmds, delta)                             # clamp, __update_angle, __update_position
            self._pull(theta_before, last_before, write_stats)
            return
        ctrl0 = self.robots[0].controller
        evolve = bool(getattr(ctrl0, "evolve", False))
        delta_ctrl = getattr(ctrl0, "delta_ctrl", delta)
        if noisy:
            sim.set_noise("off")
            clean = sim.sense()                               # fixes every robot's draw count (see the module docstring)
            sim.set_noise("replay", stream=self._noise_streams(clean))
        else:
            sim.set_noise("off")
        sim.run(1, delta=delta, delta_ctrl=delta_ctrl, evolve=evolve)
        self._pull(theta_before, last_before, write_stats)
