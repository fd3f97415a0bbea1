"""Unit scaling (farms_core ``SimulationUnitScaling`` stand-in).

The reference authors the MJCF in scaled units and divides logs back to SI
(farms_mujoco/simulation/physics.py:428-523, mjcf.py:169,567,825,840,1338,
1432-1433, swimming/drag.pyx:342-344,363,372).  Only the attributes the hot
path touches are provided.
"""


class SimulationUnitScaling:
    """Derived unit factors from (meters, seconds, kilograms)."""

    def __init__(self, meters=1.0, seconds=1.0, kilograms=1.0):
        self.meters = float(meters)
        self.seconds = float(seconds)
        self.kilograms = float(kilograms)

    @property
    def velocity(self):
        return self.meters/self.seconds

    @property
    def angular_velocity(self):
        return 1.0/self.seconds

    @property
    def acceleration(self):
        return self.meters/self.seconds**2

    @property
    def newtons(self):
        return self.kilograms*self.acceleration

    @property
    def torques(self):
        return self.kilograms*self.meters**2/self.seconds**2

    @property
    def inertia(self):
        return self.kilograms*self.meters**2

    @property
    def angular_stiffness(self):
        return self.torques

    @property
    def angular_damping(self):
        return self.torques*self.seconds

    def as_tuple(self):
        return (self.meters, self.seconds, self.kilograms)
