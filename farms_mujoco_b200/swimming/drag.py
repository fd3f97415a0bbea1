"""Swimming (mirror of farms_mujoco/swimming/drag.pyx).

The resistive-force model (``drag_forces`` drag.pyx:152-268, ``compute_buoyancy``
:111-149) and the downstream applier that writes ``xfrc_applied`` are
fused into the CUDA step: per swimming link, every step, for every environment, without
leaving the SM.  The classes below keep the reference's construction / stepping protocol
and forward the knobs to the engine.  ``drag_forces`` below is the reference's operator of that
name for callers that hold link rows of their own: it runs on the device too (``fb_drag_forces``,
float64); there is no host arithmetic.
"""

import numpy as np


def drag_forces(iteration, data_links, links_index, data_xfrc, xfrc_index, coefficients, z3, z4,
                water, mass, height, density, gravity, use_buoyancy):
    """Drag swimming (drag.pyx:152-268, same arguments; ``z3`` / ``z4`` are the reference's
    scratch arrays and are not used).  Forces and torques of link ``links_index`` at ``iteration``
    are stored into ``data_xfrc.array[iteration, xfrc_index]`` in the CoM frame; returns whether
    the link was at or below the water surface.  One row through ``fb_drag_forces``; for many rows
    at once see ``engine.drag_forces_rows``."""
    # pylint: disable=too-many-arguments
    del z3, z4
    from ..engine import drag_forces_rows  # pylint: disable=import-outside-toplevel
    row = np.array(data_links.array[iteration, links_index], dtype=np.float64).reshape(1, 20)
    out = np.array(data_xfrc.array[iteration, xfrc_index, 0:6], dtype=np.float64).reshape(1, 6)
    pos = row[0, 0:3]
    applied = drag_forces_rows(
        row, np.asarray(coefficients, dtype=np.float64).reshape(1, 6), mass, height, density,
        water.surface(pos[0], pos[1]), water.velocity(pos[0], pos[1], pos[2]),
        water.viscosity(pos[0], pos[1], pos[2]), gravity, use_buoyancy, out)
    if applied[0]:
        data_xfrc.array[iteration, xfrc_index, 0:6] = out[0]
    return bool(applied[0])


class WaterProperties:
    """Water properties (drag.pyx:271-306)"""

    def __init__(self, surface, density, velocity, viscosity):
        self._surface = float(surface)
        self._density = float(density)
        self._velocity = np.array(velocity, dtype=float)
        self._viscosity = float(viscosity)

    def surface(self, x=0.0, y=0.0):
        return self._surface

    def density(self, x=0.0, y=0.0, z=0.0):
        return self._density

    def velocity(self, x=0.0, y=0.0, z=0.0):
        return self._velocity

    def viscosity(self, x=0.0, y=0.0, z=0.0):
        return self._viscosity

    def set_velocity(self, vx, vy, vz):
        self._velocity[:] = (vx, vy, vz)


class SwimmingHandler:
    """Swimming handler (drag.pyx:309-419): same constructor, ``step`` and
    ``set_water_velocity``; ``physics`` is a ``BatchedPhysics`` whose tables already hold
    the per-link masses, heights, densities and coefficients (drag.pyx:353-385 ->
    FarmsTables.swim_*)."""

    def __init__(self, data, animat_options, arena_options, units, physics):
        del units
        self.animat_options = animat_options
        self.links = data.sensors.links
        self.xfrc = data.sensors.xfrc
        self.physics = physics
        water_options = arena_options.water
        self.drag = bool(water_options.drag)
        self.sph = bool(getattr(water_options, 'sph', False))
        self.buoyancy = bool(water_options.buoyancy)
        self.water = WaterProperties(
            surface=float(water_options.height), density=float(water_options.density),
            velocity=np.array(water_options.velocity, dtype=float),
            viscosity=float(water_options.viscosity))
        self.n_links = len([link for link in animat_options.morphology.links if link.swimming])
        physics.set_swimming(self.drag, self.buoyancy)
        physics.set_water_velocity(self.water.velocity())

    def step(self, iteration):
        """Swimming step: the forces of ring row ``iteration`` were computed (and applied)
        on the device; copy them to ``data.sensors.xfrc`` as the reference's step leaves
        them (drag.pyx:265-267)."""
        if (self.drag or self.sph) and self.xfrc.array.shape[2]:
            self.xfrc.array[:, iteration] = self.physics.log_row('xfrc', iteration, self.physics.log_stride > 1)

    def set_water_velocity(self, velocity):
        """Set water velocity (drag.pyx:417-419)"""
        self.water.set_velocity(velocity[0], velocity[1], velocity[2])
        self.physics.set_water_velocity(velocity)
