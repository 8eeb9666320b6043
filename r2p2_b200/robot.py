"""Robot facade -- the reference's ``Robot`` (r2p2/robot.py:43-402) over the batched CUDA step.

Same constructor arguments, attributes (``x, y, orientation, speed, angular_velocity, last_pos, sensors,
vision_range, radius, max_speed, acceleration, controller, has_noise``) and the same entry point
``update(env, delta, write_stats=False)``.  A robot belongs to a ``RobotGroup`` (r2p2_b200/group.py): alone when it is
updated on its own, together with the other robots of its scenario when ``simulation.Simulation`` drives it -- one
launch per tick per group either way.  There is no host implementation of the physics: without the CUDA library every
``update`` raises.
"""
from .controller import Controller


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
        self._group = None
        self._env_id = None
        self._genome_dirty = True
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
        """robot.py:365-371: stop dead."""
        self.speed = 0

    def brake(self):
        """robot.py:354-363: braking proportional to acceleration and speed; returns the acceleration change."""
        neg = self.speed < 0
        new_speed = self.speed - self.acceleration - self.speed / 3
        if (neg and new_speed > 0) or (not neg and new_speed < 0):
            self.stop()
            return -self.acceleration / 5
        return -self.acceleration - self.speed / 3

    def change_acceleration(self, acceleration):
        self.acceleration = acceleration

    def turn(self, angle=0):
        self.orientation += angle

    def collide(self, spot):
        """Robot.collide (robot.py:170-179): the controller's callback (the log line is written by the group)."""
        self.collisions += 1
        self.controller.on_collision(spot)

    def check_sensors(self, env):
        """(col, dst, ang) for the current pose without moving the robot (robot.py:287-318): one r2p2_sense launch."""
        import numpy as np
        from .group import ray_angles
        from .logs import python_dst
        g = self._ensure_group(env)
        g._push()
        i = g.robots.index(self)
        g.sim.set_noise("off")
        dst_all, px = g.sim.sense(hit_pixels=True)
        row = dst_all[i].copy()
        if self.has_noise:
            cf = float(g.clamp)
            for k in range(len(row)):
                if row[k] != cf:
                    row[k] += np.random.normal(0, 0.5)        # robot.py:308-310
        dst = python_dst(row, g.clamp)
        ang = [a + self.orientation for a in ray_angles(self.sensors)]
        col = [(float(p[0]), float(p[1])) if p[0] >= 0 else (-1, -1) for p in px[i]]
        return col, dst, ang

    def update(self, env, delta, write_stats=False):
        """Robot.update (robot.py:389-402): sense, control, turn, move.  A robot driven on its own is a group of one."""
        g = self._ensure_group(env)
        if len(g.robots) != 1:
            raise RuntimeError("this robot belongs to a Simulation group of %d robots: advance it with Simulation.update" % len(g.robots))
        g.update(delta, write_stats)

    # ---- plumbing --------------------------------------------------------------------------------
    def _genome_changed(self):
        self._genome_dirty = True

    def _ensure_group(self, env):
        from .group import RobotGroup
        if self._group is not None and self._env_id != id(env) and len(self._group.robots) == 1:
            self._group.close()
            self._group = None
        if self._group is None:
            RobotGroup([self], env, device=self.device)
            self._env_id = id(env)
        return self._group
