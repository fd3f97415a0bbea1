"""Controller protocol of the simulation loop (stand-in for the farms_core contract).

The reference task drives a ``farms_core.model.control.AnimatController``
(farms_mujoco/simulation/task.py:15,288-346): ``joints_names[ControlType]``,
``muscles_names``, ``step / positions / torques / springrefs / excitations
(iteration, time, timestep)``.  farms_core is not in this image, so the two
names the path touches are restated here; controllers written against
farms_core satisfy this protocol unchanged.

``TravellingWaveController`` is the batched position controller of the
benchmark workloads.  It can be evaluated on the host (reference ordering, one
``ctrl`` upload per iteration) or handed to the engine, which then evaluates it
inside the step kernel (``device_parameters``).
"""

import enum

import numpy as np


class ControlType(enum.IntEnum):
    """farms_core.model.control.ControlType (task.py:229-246)"""
    POSITION = 0
    VELOCITY = 1
    TORQUE = 2


class AnimatController:
    """Protocol base: every method of the farms_core controller the task calls."""

    def __init__(self, joints_names=None, muscles_names=()):
        self.joints_names = joints_names if joints_names is not None else [[], [], []]
        self.muscles_names = list(muscles_names)

    def step(self, iteration, time, timestep):
        """Advance the controller state (task.py:292-296)."""

    def positions(self, iteration, time, timestep):
        """{joint name: position command (scalar or [n_envs])} (task.py:312-321)."""
        return {}

    def torques(self, iteration, time, timestep):
        """{joint name: torque command} in SI units (task.py:326-337)."""
        return {}

    def springrefs(self, iteration, time, timestep):
        """{joint name: spring reference} (task.py:338-346)."""
        return {}

    def excitations(self, iteration, time, timestep):
        """Muscle excitations (task.py:299-307); muscles are outside this path."""
        return []


class TravellingWaveController(AnimatController):
    """``position_j = offset_j + A_j sin(2 pi f_j t - lag_j + phase_env)``."""

    def __init__(self, joints, amplitude, frequency, phase_lag, env_phase=0.0, offset=None):
        super().__init__(joints_names=[list(joints), [], []])
        self.amplitude = np.asarray(amplitude, dtype=float)
        self.frequency = np.asarray(frequency, dtype=float)
        self.phase_lag = np.asarray(phase_lag, dtype=float)
        self.offset = np.zeros_like(self.amplitude) if offset is None else np.asarray(offset, dtype=float)
        self.env_phase = np.atleast_1d(np.asarray(env_phase, dtype=float))

    def positions(self, iteration, time, timestep):
        phase = 2*np.pi*self.frequency[None, :]*time - self.phase_lag[None, :] + self.env_phase[:, None]
        values = self.offset[None, :] + self.amplitude[None, :]*np.sin(phase)
        return {joint: values[:, i] for i, joint in enumerate(self.joints_names[ControlType.POSITION])}

    def device_parameters(self):
        """What ``fb_set_wave_controller`` / ``fb_set_env_phase`` need (include/farms_b200.h)."""
        return dict(joints=self.joints_names[ControlType.POSITION], amplitude=self.amplitude,
                    frequency=self.frequency, phase_lag=self.phase_lag, offset=self.offset,
                    env_phase=self.env_phase)
