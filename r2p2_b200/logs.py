"""The reference's per-robot log files, written from the device trace.

Every ``Robot.update`` of the reference appends a sensor block to ``../logs/robot_<id>.log`` and one line to
``../logs/robot_<id>_sensors.log`` (r2p2/robot.py:337-352, called from ``control`` at robot.py:378), and with
``write_stats=True`` a pose snapshot (``write_stats_to_log``, robot.py:158-168).  Tools around the simulator
(sensor-log plots, ``benchmarker.py``) read these files, so a batched run can emit them after the fact from the trace
the kernel records (``BatchSim.set_trace`` / ``trace``): the text below is byte-identical to what the reference writes
for the same episode (tests/golden/logs.json), including the Python typing quirks -- a clamped reading prints the
JSON-typed ``vision_range[1] + radius`` (``55``, not ``55.0``), tick 0 prints the JSON-typed orientation.
"""
import math
import os


def python_dst(dst_row, clamp_value):
    """Readings of one tick with the reference's Python types: a clamped ray holds ``vision_range[1] + radius`` itself
    (robot.py:305-306), every other one a float.  A reading equal to the clamp value is a clamped one: the ray loop only
    runs while norm < limit, so an unclamped hit is strictly closer."""
    cf = float(clamp_value)
    return [clamp_value if float(d) == cf else float(d) for d in dst_row]


def format_sensor_data(identifier, ang, orientation, dst):
    """(text appended to the robot log, line appended to the sensors log) -- robot.py:337-352."""
    text = "[ROBOT ID " + str(identifier) + " SENSOR DATA:\n"
    for i in range(len(ang)):
        text += "SENSOR: " + str(ang[i] - orientation) + "º - DISTANCE: " + str(dst[i]) + "\n"
    text += "]"
    line = ";".join(str(d) for d in dst) + "\n"
    return text, line


def format_stats(identifier, x, y, orientation, speed, max_speed, acceleration):
    """``write_stats_to_log`` snapshot (robot.py:164-166)."""
    return ("\n[ROBOT ID: " + str(identifier) + "; POSITION: (" + str(x) + ", " + str(y) + "); ORIENTATION: " + str(orientation) +
            "º; SPEED: " + str(speed) + "; MAX_SPEED: " + str(max_speed) + "; ACCELERATION: " + str(acceleration) + "]\n")


def format_collision(identifier, spot):
    """``Robot.collide`` (robot.py:177): no closing bracket, no newline -- as the reference writes it."""
    return "[ROBOT ID " + str(identifier) + " COLLISION AT (" + str(spot[0]) + ", " + str(spot[1]) + ")"


def collision_spots(cinfo_row, shape, last_pos):
    """The ``collide(spot)`` calls of one tick, in the reference's order (robot.py:211-253), from a row of the
    collision trace (``r2p2_read_trace_collisions``): wall contact at the proposed position, one of the eight map-border
    cases, the crash-radius test (spot = the position the robot is put back to)."""
    code = int(cinfo_row[0])
    W, H = int(shape[0]), int(shape[1])
    spots = []
    if code & 1:
        spots.append((float(cinfo_row[1]), float(cinfo_row[2])))
    b = float(cinfo_row[3])
    case = (code >> 1) & 15
    if case:
        spots.append({1: (W, b), 2: (W, H), 3: (W, 0), 4: (0, b), 5: (0, H), 6: (0, 0), 7: (b, H), 8: (b, 0)}[case])
    if code & 32:
        spots.append(tuple(last_pos))
    return spots


def format_controller_step(kind, odom, dst, out, vision_range1, stamp):
    """One ``log_step`` line of the controller's own log -- ``../logs/neuro.log`` (neurocontroller.py:94-98) or
    ``../logs/inspyred.log`` (inspyred.py:97-103): ``[time]\tOdom: (x, y)\t-\tIn: ...\t-\tOut: (a, b)``.

    The reference prints Python objects, so the types matter: an unclamped distance is a numpy float (``np.linalg.norm``),
    a clamped one the JSON-typed range; the controller's answer is a tuple of numpy floats; ``odom`` is whatever the
    robot's ``x, y`` are at that moment (the caller passes them typed).  kind "neuro": In is the normalised input
    ``vision_range[1] / dst`` as a 1 x R numpy array (printed by numpy, 8 digits); kind "linear": In is the distance list
    itself.  ``stamp`` stands where the reference puts ``time.time()``."""
    import numpy as np
    o = (np.float64(out[0]), np.float64(out[1]))
    if kind == "neuro":
        with np.errstate(divide="ignore"):
            inpt = np.asarray([vision_range1 / np.float64(d) for d in dst]).reshape(1, -1)
    else:
        inpt = [d if not isinstance(d, float) else np.float64(d) for d in dst]      # python_dst: floats are the unclamped readings
    return "[" + str(stamp) + "]\tOdom: " + str(odom) + "\t-\tIn: " + str(inpt) + "\t-\tOut: " + str(o) + "\n"


def episode_logs(pose, dst, *, identifier, sensors, orientation0, vision_range, radius, max_speed, acceleration=0,
                 write_stats=True, cinfo=None, shape=None, position0=None, controller=None, ticks=None, stamps=None):
    """Log texts of one robot's episode.  pose [T, 5] = (x, y, theta, speed, omega) after each tick, dst [T, R] and
    (optionally, for the COLLISION lines) cinfo [T, 4] as ``BatchSim.trace`` returns them for that robot, shape = the
    stage's (W, H), position0 = the JSON-typed start position; the remaining arguments carry the robot JSON's own
    (Python-typed) values.  returns {"log": ..., "sensors": ...}.

    controller = "neuro" | "linear" (needs the six-column cinfo, whose last two columns are the controller's answer, and
    position0): also returns "controller", the text of ``neuro.log`` / ``inspyred.log`` -- one line per tick in which
    ``control()`` reached its ``log_step`` (``ticks``: how many that were, default all), stamped with ``stamps[t]`` (default
    ``time.time()`` when the text is built: the reference stamps wall-clock time, the one field that cannot be reproduced)."""
    import time as _time
    # the ray angles with the types the reference iterates over: the JSON list itself, or range() ints (utils.py:465-466)
    rays = list(sensors) if isinstance(sensors, (list, tuple)) else list(range(0, 360, math.floor(365 / int(sensors))))
    clamp = vision_range[1] + radius
    log, sens = [], []
    theta = orientation0
    last_pos = tuple(position0) if position0 is not None else None   # Robot.last_pos keeps the JSON types until the first free move
    ctl = []
    odom = tuple(position0) if position0 is not None else None       # Controller.odom = (robot.x, robot.y) with the robot's own types
    n_ctl = len(pose) if ticks is None else int(ticks)
    # numpy-ness of x and y: the JSON pose is Python-typed; a move adds cos * speed * delta with the controller's numpy
    # speed (robot.py:196-197), the border code assigns Python numbers (robot.py:222-247), a crash restores last_pos
    npx = npy = last_npx = last_npy = False
    if controller is not None:
        import numpy as _np
    for t in range(len(pose)):
        ang = [a + theta for a in rays]                     # utils.py:467 / 505: angles.append(a + offset)
        pd = python_dst(dst[t], clamp)
        text, line = format_sensor_data(identifier, ang, theta, pd)
        log.append(text)
        sens.append(line)
        if controller is not None and t < n_ctl:
            stamp = stamps[t] if stamps is not None else _time.time()
            ctl.append(format_controller_step(controller, odom, pd, (cinfo[t][4], cinfo[t][5]), vision_range[1], stamp))
        x, y, theta, speed = (float(v) for v in pose[t][:4])
        if cinfo is not None:
            crashed = int(cinfo[t][0]) & 32
            for spot in collision_spots(cinfo[t], shape, last_pos if last_pos is not None else (x, y)):
                log.append(format_collision(identifier, spot))
            case = (int(cinfo[t][0]) >> 1) & 15
            if case:                                        # the border code assigns Python ints (robot.py:222-247)
                W, H = int(shape[0]), int(shape[1])
                if case <= 3: x = W - 1
                elif case <= 6: x = 0
                if case in (2, 5, 7): y = H - 1 - radius
                elif case in (3, 6, 8): y = radius
            npx, npy = True, True
            if case:
                if case <= 6: npx = False
                if case in (2, 3, 5, 6, 7, 8): npy = False
            if crashed and last_pos is not None:
                x, y = last_pos                             # self.x = self.last_pos[0] (robot.py:250-251)
                npx, npy = last_npx, last_npy
            else:
                last_pos = (x, y)
                last_npx, last_npy = npx, npy
        else:
            npx = npy = True
        if controller is not None:
            odom = (_np.float64(x) if npx else x, _np.float64(y) if npy else y)
        if write_stats:
            log.append(format_stats(identifier, x, y, theta, speed, max_speed, acceleration))
    out = {"log": "".join(log), "sensors": "".join(sens)}
    if controller is not None:
        out["controller"] = "".join(ctl)
    return out


class RobotLogs:
    """The three files the reference's ``Robot.__init__`` opens (robot.py:84-93); used by the ``Robot`` facade when it
    is given a ``log_dir``."""

    def __init__(self, log_dir, identifier, name=None):
        stem = str(name) if name is not None else "robot_" + str(identifier)
        os.makedirs(log_dir, exist_ok=True)
        self.log = open(os.path.join(log_dir, stem + ".log"), "w+")
        self.sensors_output = open(os.path.join(log_dir, stem + "_sensors.log"), "w+")
        self.benchmark_log = open(os.path.join(log_dir, stem + "_benchmark.log"), "w+")

    def sensor_data(self, identifier, ang, orientation, dst):
        text, line = format_sensor_data(identifier, ang, orientation, dst)
        self.log.write(text)
        self.sensors_output.write(line)

    def stats(self, robot):
        self.log.write(format_stats(robot.identifier, robot.x, robot.y, robot.orientation, robot.speed, robot.max_speed, robot.acceleration))

    def close(self):
        for f in (self.log, self.sensors_output, self.benchmark_log):
            f.close()
