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

``CPGController`` is a network of coupled phase oscillators with amplitude dynamics (the
salamander-type CPG FARMS controllers implement in farms_core / farms_amphibious) driving
position and/or torque actuators.  On the host it is the NumPy reference, stepped through
``ExperimentTask.step_control`` exactly like any farms_core controller; ``device_cpg`` hands the
same network to the engine (``fb_set_cpg``), which integrates it inside every launch.
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


class CPGController(AnimatController):
    """Coupled phase oscillators (explicit Euler at the physics time step):

        theta_i' = 2 pi f_i + sum_j w_ij r_j sin(theta_j - theta_i - phi_ij)
        r_i''    = a_i (a_i/4 (R_i - r_i) - r_i')

    ``outputs``: list of ``(joint, ControlType, osc_a, osc_b, gain, offset)``; the command of the
    joint is ``offset + gain (r_a (1 + cos theta_a) - r_b (1 + cos theta_b))`` (``osc_b = -1``:
    ``offset + gain r_a cos theta_a``), a position target or a torque in SI units.
    ``couplings``: list of ``(from_j, to_i, w_ij, phi_ij)``.  State arrays are ``[n_envs, n_osc]``.
    ``springs``: list of ``(joint, osc_a, osc_b, gain, offset)``: spring references of passive
    joints driven by the same expression (``springrefs``; the reference applies them together with the
    torque commands, task.py:338-346).
    The command of iteration k is the output after k Euler steps: ``step`` advances the state
    when the task moves to a new iteration (task.py:292-296)."""
    # pylint: disable=too-many-instance-attributes,too-many-arguments

    def __init__(self, frequency, amplitude, rate, couplings, outputs, phase0, amplitude0=None, device=True, springs=()):
        pos = [o[0] for o in outputs if ControlType(o[1]) == ControlType.POSITION]
        trq = [o[0] for o in outputs if ControlType(o[1]) == ControlType.TORQUE]
        super().__init__(joints_names=[pos, [], trq])
        self.frequency = np.asarray(frequency, dtype=float)
        self.amplitude = np.asarray(amplitude, dtype=float)
        self.rate = np.asarray(rate, dtype=float)
        self.couplings = [(int(j), int(i), float(w), float(phi)) for j, i, w, phi in couplings]
        self.outputs = [(str(j), ControlType(t), int(a), int(b), float(g), float(o)) for j, t, a, b, g, o in outputs]
        self.springs = [(str(j), None, int(a), int(b), float(g), float(o)) for j, a, b, g, o in springs]
        self.theta = np.array(np.atleast_2d(phase0), dtype=float)
        self.r = np.zeros_like(self.theta) if amplitude0 is None else np.array(
            np.broadcast_to(amplitude0, self.theta.shape), dtype=float)
        self.rd = np.zeros_like(self.theta)
        self._state0 = (self.theta.copy(), self.r.copy())
        self._last = None
        self._device = bool(device)
        if not device:
            # instance attribute shadowing the method: the task then evaluates it on the host
            self.device_cpg = None

    def _advance(self, timestep):
        dth = np.tile(2*np.pi*self.frequency, (self.theta.shape[0], 1))
        for j, i, w, phi in self.couplings:
            dth[:, i] += w*self.r[:, j]*np.sin(self.theta[:, j] - self.theta[:, i] - phi)
        rdd = self.rate*(0.25*self.rate*(self.amplitude - self.r) - self.rd)
        self.theta = self.theta + timestep*dth
        self.r, self.rd = self.r + timestep*self.rd, self.rd + timestep*rdd

    def step(self, iteration, time, timestep):
        key = int(round(time/timestep)) if timestep else iteration
        if self._last is not None and key != self._last:
            for _ in range(key - self._last):
                self._advance(timestep)
        self._last = key

    def _command(self, out):
        _, _, a, b, gain, offset = out
        if b >= 0:
            val = self.r[:, a]*(1 + np.cos(self.theta[:, a])) - self.r[:, b]*(1 + np.cos(self.theta[:, b]))
        else:
            val = self.r[:, a]*np.cos(self.theta[:, a])
        return offset + gain*val

    def positions(self, iteration, time, timestep):
        return {o[0]: self._command(o) for o in self.outputs if o[1] == ControlType.POSITION}

    def torques(self, iteration, time, timestep):
        return {o[0]: self._command(o) for o in self.outputs if o[1] == ControlType.TORQUE}

    def springrefs(self, iteration, time, timestep):
        return {s[0]: self._command(s) for s in self.springs}

    def device_cpg(self, actuator_index, torque_unit=1.0, spring_index=None):  # pylint: disable=method-hidden
        """The network as ``fb_set_cpg`` takes it.  ``actuator_index(joint, ControlType)`` maps
        an output to its ctrl index; torque commands are scaled by ``units.torques`` (task.py:332);
        ``spring_index(joint)`` maps a spring reference to its qpos address."""
        theta0, r0 = self._state0
        springs = self.springs if spring_index is not None else []
        return dict(
            spring_qpos_adr=[spring_index(s[0]) for s in springs],
            spring_osc_a=[s[2] for s in springs], spring_osc_b=[s[3] for s in springs],
            spring_gain=[s[4] for s in springs], spring_offset=[s[5] for s in springs],
            frequency=self.frequency, amplitude=self.amplitude, rate=self.rate,
            coupling_from=[c[0] for c in self.couplings], coupling_to=[c[1] for c in self.couplings],
            coupling_weight=[c[2] for c in self.couplings], coupling_bias=[c[3] for c in self.couplings],
            out_actuator=[actuator_index(o[0], o[1]) for o in self.outputs],
            out_osc_a=[o[2] for o in self.outputs], out_osc_b=[o[3] for o in self.outputs],
            out_gain=[o[4]*(torque_unit if o[1] == ControlType.TORQUE else 1.0) for o in self.outputs],
            out_offset=[o[5]*(torque_unit if o[1] == ControlType.TORQUE else 1.0) for o in self.outputs],
            phase0=theta0, amplitude0=r0)
