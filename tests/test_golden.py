"""Committed golden vectors (tests/golden/*.npz, made by make_golden.py from the
oracle): the oracle must still reproduce them bit-for-bit-ish (CPU), the device
code must match them within the fp32 tolerance (emulation on CPU, CUDA on GPU)."""

import os

import numpy as np
import pytest

from conftest import make_case, oracle_rollout, scaled_error, log_error, state_errors

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = [('swimmer8', 5e-4), ('salamander_swim', 5e-4), ('salamander', 5e-3), ('centipede', 5e-3)]


def _load(name):
    return np.load(os.path.join(GOLDEN, f'{name}.npz'))


@pytest.mark.parametrize('name,_tol', CASES)
def test_oracle_reproduces_golden(name, _tol):
    gold = _load(name)
    spec, model, qpos0, qvel0, ctrl = make_case(name, 2, seed=7)
    assert np.array_equal(qpos0, gold['qpos0']) and np.array_equal(ctrl, gold['ctrl'])
    from farms_mujoco_b200.data import AnimatData
    from farms_mujoco_b200.simulation.physics import FarmsTables
    from oracle import farms_oracle as fo
    data = AnimatData.from_sensors_names(model.timestep, 1, spec.links_names, spec.joints_names,
                                         spec.contacts_names, spec.xfrc_names)
    maps = fo.make_maps(model, data)
    tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options,
                         spec.arena_options, spec.simulation_options.units)
    n_rows = int(gold['n_rows'])
    for env in range(2):
        _, log, states = oracle_rollout(spec, model, tables, n_rows, qpos0[env], qvel0[env], ctrl[env])
        assert np.allclose(states[-1][0], gold[f'qpos_{env}'], rtol=1e-9, atol=1e-11)
        assert np.allclose(states[-1][1], gold[f'qvel_{env}'], rtol=1e-8, atol=1e-10)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.allclose(getattr(log.sensors, kind).array, gold[f'{kind}_{env}'],
                               rtol=1e-7, atol=1e-9), kind


def _engine_vs_golden(library, name, tol, **kw):
    from farms_mujoco_b200.engine import BatchedPhysics
    gold = _load(name)
    spec, model, _, _, _ = make_case(name, 2, seed=7)
    n_rows = int(gold['n_rows'])
    physics = BatchedPhysics.from_spec(spec, 2, buffer_size=n_rows, library=library, **kw)
    physics.reset(gold['qpos0'], gold['qvel0'])
    physics.set_ctrl(gold['ctrl'])
    physics.step(n_rows - 1)
    logs = physics.log_arrays()
    for env in range(2):
        errs = state_errors(physics.qpos[env], physics.qvel[env], gold[f'qpos_{env}'], gold[f'qvel_{env}'])
        assert max(errs.values()) < tol, errs
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert log_error(kind, logs[kind][env], gold[f'{kind}_{env}']) < tol, kind


@pytest.mark.parametrize('name,tol', CASES)
def test_emulated_device_code_matches_golden(emu_library, name, tol):
    _engine_vs_golden(emu_library, name, tol)


@pytest.mark.gpu
@pytest.mark.parametrize('name,tol', CASES)
def test_cuda_matches_golden(cuda_library, name, tol):
    _engine_vs_golden(cuda_library, name, tol)
