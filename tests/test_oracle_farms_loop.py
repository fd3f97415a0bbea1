"""oracle/farms_loop.c (the CPU baseline's compiled loop) against oracle/farms_oracle.py (the
NumPy restatement of farms_mujoco/simulation/physics.py:435-545, sensors/sensors.pyx:140-190 and
swimming/drag.pyx:152-268): same fp64 log, row for row."""

import numpy as np
import pytest

from farms_mujoco_b200 import models, mjcf_subset
from farms_mujoco_b200.data import AnimatData
from farms_mujoco_b200.models import travelling_wave_parameters
from farms_mujoco_b200.simulation.physics import FarmsTables
from oracle.oracle import OraclePhysics
from oracle import farms_oracle as fo


def _setup(name):
    spec = models.MODELS[name]()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    data = AnimatData.from_sensors_names(model.timestep, 4, spec.links_names, spec.joints_names,
                                         spec.contacts_names, spec.xfrc_names)
    maps = fo.make_maps(model, data)
    tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options, spec.arena_options,
                         spec.simulation_options.units)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    acts = np.array([model.actuator_id(f'actuator_position_{j}') for j in joints])
    return spec, model, tables, (acts, amp, freq, lag)


@pytest.mark.parametrize('name', ['swimmer8', 'salamander_swim', 'salamander', 'centipede'])
def test_compiled_loop_equals_numpy_restatement(name):
    spec, model, tables, wave = _setup(name)
    acts, amp, freq, lag = wave
    n_it, phase = 40, 0.7
    rng = np.random.default_rng(3)
    qpos0 = np.array(model.key_qpos, dtype=float)
    qpos0[7:] += rng.uniform(-0.1, 0.1, model.nq - 7)
    qvel0 = rng.uniform(-0.2, 0.2, model.nv)

    def controller(iteration, time):
        ctrl = np.zeros(model.nu)
        ctrl[acts] = amp*np.sin(2*np.pi*freq*time - lag + phase)
        return ctrl
    ref, states = fo.reference_rollout(OraclePhysics(model), spec, tables, n_it, controller=controller,
                                       qpos0=qpos0, qvel0=qvel0)
    physics = OraclePhysics(model)
    physics.reset(keyframe_id=0)
    physics.data.qpos[:] = qpos0
    physics.data.qvel[:] = qvel0
    physics.forward()
    loop = fo.CompiledRollout(physics, spec, tables, n_it, wave=wave, env_phase=phase)
    loop.run(17, timed=True)
    loop.run(n_it - 1 - 17)
    loop.sensors()
    assert np.allclose(physics.data.qpos, states[-1][0], rtol=0, atol=1e-13)
    assert np.allclose(physics.data.qvel, states[-1][1], rtol=1e-12, atol=1e-12)
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        ours, want = getattr(loop.data.sensors, kind).array, getattr(ref.sensors, kind).array
        assert ours.shape == want.shape
        assert np.allclose(ours, want, rtol=1e-11, atol=1e-12), (kind, np.abs(ours - want).max())
    if name in ('salamander', 'centipede'):
        assert np.abs(ref.sensors.contacts.array).max() > 0
    else:
        assert np.abs(ref.sensors.xfrc.array).max() > 0
    assert loop.stage_seconds[3] > 0 and loop.stage_seconds.sum() < 5.0


def test_compiled_loop_ring_reuse():
    """buffer_size < iterations: rows are zeroed before they are revisited (the `+=` columns)."""
    spec, model, tables, wave = _setup('salamander')
    physics = OraclePhysics(model)
    physics.reset(keyframe_id=0)
    loop = fo.CompiledRollout(physics, spec, tables, 8, wave=wave)
    loop.run(8)
    first = loop.data.sensors.contacts.array.copy()
    physics.reset(keyframe_id=0)
    loop.iteration = 8
    loop.run(8)
    # same state sequence written over the same ring rows, time-dependent control aside
    assert np.abs(loop.data.sensors.contacts.array).max() < 2*np.abs(first).max() + 1e-12


def test_compiled_loop_with_pair_contacts():
    """The same on explicit <pair> self-collisions: the pair sensor (g1, g2) and the two single-link
    sensors (g, -1) of sensors.pyx:160-176 see the two-body contact in the C loop as in NumPy."""
    import variant_models
    spec = variant_models.salamander_foot_pairs()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    data = AnimatData.from_sensors_names(model.timestep, 4, spec.links_names, spec.joints_names,
                                         spec.contacts_names, spec.xfrc_names)
    maps = fo.make_maps(model, data)
    tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options, spec.arena_options,
                         spec.simulation_options.units)
    n_it = 12
    qpos0 = variant_models.folded_legs_qpos(model, 1.0)
    ref, states = fo.reference_rollout(OraclePhysics(model), spec, tables, n_it, qpos0=qpos0, qvel0=np.zeros(model.nv))
    physics = OraclePhysics(model)
    physics.reset(keyframe_id=0)
    physics.data.qpos[:] = qpos0
    physics.forward()
    loop = fo.CompiledRollout(physics, spec, tables, n_it)
    loop.run(n_it - 1)
    loop.sensors()
    assert np.allclose(physics.data.qpos, states[-1][0], rtol=0, atol=1e-13)
    ours, want = loop.data.sensors.contacts.array, ref.sensors.contacts.array
    names = [tuple(c) for c in spec.contacts_names]
    pair = names.index(('link_leg_0_L_3', 'link_leg_0_R_3'))
    assert np.abs(want[1:, pair, 6:9]).max() > 1e-3
    assert np.allclose(ours, want, rtol=1e-11, atol=1e-12), np.abs(ours - want).max()
