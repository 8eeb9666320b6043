"""Naive_controller -- constant command, as the reference's r2p2/controllers/naive_controller.py:35-55 (returns (3, 0))."""
from ..controller import Controller


class Naive_controller(Controller):
    device_kind = "const"

    def __init__(self, config=None, speed=3.0, angular_velocity=0.0):
        super().__init__("NAIVE", config)
        self.command = (float(speed), float(angular_velocity))

    def control(self, dst):
        self.update_sensor_angles(self.ang, dst)
        return self.command

    def device_spec(self, n_rays):
        return dict(kind="const", hidden_layer=None, activation="tanh", genome=list(self.command))


Naive_Controller = Naive_controller      # the reference's spelling (conf/controller-naive.json: "naive_controller.Naive_Controller")
