"""Robot facade -- the reference's ``Robot`` (r2p2/robot.py:43-402) over the batched CUDA step.

Same constructor arguments, attributes (``x, y, orientation, speed, angular_velocity, last_pos, sensors,
vision_range, radius, max_speed, controller, has_noise``) and the same entry point
``update(env, delta, write_stats=False)``.  One ``update`` is one launch of the fused step kernel for this
robot (P = 1); lists of robots that share a stage should go through ``simulation.Simulation`` which batches
them into one launch.  There is no host implementation of the physics: without the CUDA library every
``update`` raises.
"""
import numpy as np

from .controller import Controller
from .sim import BatchSim


class Robot:
    def __init__(self, identifier=-1, x=0, y=0, orientation=0, speed=0, max_speed=2, acceleration=0, sensors=8,
                 vision_range=(0, 10000), radius=1, color=None, controller=None, name=None, lock=None, device=0, log_dir=None):
        self.identifier = identifier
        self.x, self.y = x, y
        self.last_pos = (x, y)
        self.orientation = orientation
        self.speed = speed
        self.max_speed = max_speed
        self.acceleration = acceleration
        self.angular_velocity = 0
        self.sensors = sensors
        self.vision_range = vision_range
        self.radius = radius
        self.controller = controller if controller is not None else Controller()
        self.color = color
        self.name = name
        self.has_noise = True           # the reference's default (robot.py:96)
        self.device = device
        self.collisions = 0
        self.flags = 0
        self._sim = None
        self._env_id = None
        self._mirror = None
        self._dirty = True
        # the reference always writes ../logs/robot_<id>.log and ..._sensors.log (robot.py:84-93); here only on request
        self._logs = None
        if log_dir is not None:
            from .logs import RobotLogs
            self._logs = RobotLogs(log_dir, identifier, name)
        self.controller.register_robot(self)

    # ---- reference API -----------------------------------------------------------------------
    def set_position(self, x, y):
        self.x, self.y = x, y

    def set_last_position(self, x, y):
        self.last_pos = (x, y)

    def stop(self):
        self.speed = 0

    def collide(self, spot):
        self.collisions += 1
        self.controller.on_collision(spot)

    def check_sensors(self, env):
        """(col, dst, ang) for the current pose without moving the robot (robot.py:287-318).  Implemented as a
        zero-length tick with a zero command, which leaves the pose untouched; the device robot is restarted
        afterwards (odometry of a running episode is lost)."""
        sim = self._ensure(env)
        sim.set_controller("const")
        sim.upload_genomes(np.zeros((1, 2)))
        sim.reset(float(self.x), float(self.y), float(self.orientation))
        sim.set_trace(True, 1)
        sim.run(1, delta=0.0)
        tr = sim.trace(1)
        self._dirty = True
        dst = [float(d) for d in tr["dst"][0, 0]]
        ang = [a + self.orientation for a in self._ray_angles(sim)]
        col = [(float(p[0]), float(p[1])) if p[0] >= 0 else (-1, -1) for p in tr["hitpx"][0, 0]]
        return col, dst, ang

    def update(self, env, delta, write_stats=False):
        """Robot.update (robot.py:389-402): sense, control, turn, move -- one fused kernel launch."""
        sim = self._ensure(env)
        if self.controller.device_spec(sim.R) is None:
            return self._update_host_controller(sim, delta, write_stats)
        self._push(sim)
        ctrl = self.controller
        evolve = bool(getattr(ctrl, "evolve", False))
        if self.has_noise:
            # robot.py:139-140, 308-310: np.random.normal(0, 0.5) per unclamped ray, from numpy's global stream.  The tick
            # gets the next R draws of that stream to replay; afterwards the global generator is advanced by exactly the
            # number the tick consumed, so a seeded run reproduces the reference draw for draw.
            state = np.random.get_state()
            draws = np.random.normal(0, 0.5, size=sim.R)
            sim.set_noise("replay", stream=draws[None, :])
            sim.run(1, delta=delta, delta_ctrl=getattr(ctrl, "delta_ctrl", delta), evolve=evolve)
            used = int(sim.noise_draws()[0])
            np.random.set_state(state)
            if used:
                np.random.normal(0, 0.5, size=used)
        else:
            sim.set_noise("off")
            sim.run(1, delta=delta, delta_ctrl=getattr(ctrl, "delta_ctrl", delta), evolve=evolve)
        self._pull(sim)
        if write_stats and self._logs is not None:
            self._logs.stats(self)                          # write_stats_to_log (robot.py:158-168)

    def _update_host_controller(self, sim, delta, write_stats):
        """A controller without a device form (PID, PDDL, teleop, user classes): sensors and motion on the device,
        ``controller.control(dst)`` on the host in between -- Robot.control / update of the reference (robot.py:373-402)
        with the two physics halves as launches (r2p2_sense, r2p2_move)."""
        cur = (self.x, self.y, self.orientation)
        if self._dirty or self._mirror != cur or not getattr(self, "_host_mode", False):
            sim.set_controller("const")
            sim.upload_genomes(np.zeros((1, 2)))
            sim.reset(float(self.x), float(self.y), float(self.orientation))
            sim.set_trace(True, 1)
            self._dirty, self._host_mode = False, True
        if self.has_noise:
            state = np.random.get_state()
            draws = np.random.normal(0, 0.5, size=sim.R)
            sim.set_noise("replay", stream=draws[None, :])
        else:
            sim.set_noise("off")
        dst_row, px = sim.sense(hit_pixels=True)
        if self.has_noise:                                  # advance numpy's global stream by what the sweep consumed
            used = int(sim.noise_draws()[0])
            np.random.set_state(state)
            if used:
                np.random.normal(0, 0.5, size=used)
        theta_before = self.orientation
        clamp = self.vision_range[1] + self.radius
        dst = [clamp if float(d) == float(clamp) else float(d) for d in dst_row[0]]
        ang = [a + theta_before for a in self._ray_angles(sim)]
        col = [(float(p[0]), float(p[1])) for p in px[0] if p[0] >= 0]
        ctrl = self.controller
        if self._logs is not None:
            self._logs.sensor_data(self.identifier, ang, theta_before, dst)
        ctrl.handle_collision(col) if hasattr(ctrl, "handle_collision") else None
        ctrl.set_dst(dst)
        ctrl.set_ang(ang)
        speed, angular_velocity = ctrl.control(dst)         # robot.py:385
        sim.move(np.asarray([[float(speed), float(angular_velocity)]]), delta)
        s = sim.state()
        tr = sim.trace(1)
        last_before = self.last_pos
        self.x, self.y, self.orientation = float(s["x"][0]), float(s["y"][0]), float(s["theta"][0])
        self.speed, self.angular_velocity = float(s["speed"][0]), float(s["omega"][0])
        self.flags = int(s["flags"][0])
        self._mirror = (self.x, self.y, self.orientation)
        if self._logs is not None:
            from .logs import collision_spots, format_collision
            for spot in collision_spots(tr["cinfo"][0, 0], (sim.W, sim.H), last_before):
                self._logs.log.write(format_collision(self.identifier, spot))
        ncoll = int(tr["ncoll"][0, 0])
        for _ in range(ncoll):
            self.collide((self.x, self.y))
        if not ncoll:
            self.last_pos = (self.x, self.y)
        if write_stats and self._logs is not None:
            self._logs.stats(self)
        if self.flags & 2:
            raise IndexError("robot %s left the map or has a non-finite pose (the reference raises here too)" % self.identifier)

    # ---- plumbing --------------------------------------------------------------------------------
    def _ray_angles(self, sim):
        """Ray angles with the reference's own Python types (JSON list entries / range() ints, utils.py:465-466)."""
        import math
        if isinstance(self.sensors, (list, tuple)):
            return list(self.sensors)
        return list(range(0, 360, math.floor(365 / int(self.sensors))))

    def _genome_changed(self):
        self._dirty = True
        self._genome_dirty = True

    def _configure_controller(self, sim):
        spec = self.controller.device_spec(sim.R)
        self._host_mode = False
        sim.set_controller(spec["kind"], hidden_layer=spec["hidden_layer"], activation=spec["activation"])
        sim.upload_genomes(np.asarray([spec["genome"]], dtype=np.float64))
        self._genome_dirty = False

    def _ensure(self, env):
        if self._sim is None:
            self._sim = BatchSim(self.device)
            self._sim.set_robot(sonar_range=self.vision_range, radius=self.radius, max_speed=self.max_speed, sonars=self.sensors)
        if self._env_id != id(env):
            self._sim.set_stage(np.asarray(env))
            self._env_id = id(env)
            self._dirty = True
        return self._sim

    def _push(self, sim):
        """(Re)start the device robot whenever the host-side pose or genome was changed by the caller."""
        cur = (self.x, self.y, self.orientation)
        if self._dirty or getattr(self, "_genome_dirty", True) or self._mirror != cur:
            self._configure_controller(sim)
            budget = float(getattr(self.controller, "time", 0) or 0)
            sim.reset(float(self.x), float(self.y), float(self.orientation), time_budget=budget)
            sim.set_trace(True, 1)
            self._dirty = False

    def _pull(self, sim):
        s = sim.state()
        tr = sim.trace(1)
        theta_before = self.orientation
        last_before = self.last_pos
        self.x, self.y, self.orientation = float(s["x"][0]), float(s["y"][0]), float(s["theta"][0])
        self.speed, self.angular_velocity = float(s["speed"][0]), float(s["omega"][0])
        self.flags = int(s["flags"][0])
        self._mirror = (self.x, self.y, self.orientation)
        ctrl = self.controller
        dst = [float(d) for d in tr["dst"][0, 0]]
        ctrl.set_dst(dst)
        ctrl.set_ang([a + theta_before for a in self._ray_angles(sim)])
        ctrl.update_sensor_angles(ctrl.ang, list(dst))
        ctrl._fitness = float(sim.fitness()[0])
        if self._logs is not None:
            from .logs import python_dst, collision_spots, format_collision
            self._logs.sensor_data(self.identifier, ctrl.ang, theta_before, python_dst(dst, self.vision_range[1] + self.radius))
            for spot in collision_spots(tr["cinfo"][0, 0], (sim.W, sim.H), last_before):
                self._logs.log.write(format_collision(self.identifier, spot))
        ncoll = int(tr["ncoll"][0, 0])
        for _ in range(ncoll):
            self.collide((self.x, self.y))
        if not ncoll:
            self.last_pos = (self.x, self.y)
        if self.flags & 2:
            raise IndexError("robot %s left the map or has a non-finite pose (the reference raises here too)" % self.identifier)
