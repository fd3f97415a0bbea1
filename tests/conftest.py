"""Shared fixtures.  ``-m gpu`` tests need a B200; everything else runs on CPU."""

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_DIR = os.path.join(ROOT, 'tests', 'emu')
EMU_LIB = os.path.join(EMU_DIR, 'libfb_emu.so')
CSRC = os.path.join(ROOT, 'farms_mujoco_b200', 'csrc')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


@pytest.fixture(scope='session')
def emu_library():
    """Host-emulation build of the device code (unit-test harness, TEAM = 1).

    Compiles farms_mujoco_b200/csrc/fb_engine.cu with g++ -DFB_HOST_EMU so the
    fp32 arithmetic and indexing of the kernels can be checked without a GPU.
    It is test infrastructure: the product only ever loads libfarmsb200.so.
    """
    sources = [os.path.join(CSRC, f) for f in ('fb_engine.cu', 'fb_device.h', 'fb_fast.h', 'fb_fastc.h', 'fb_model.h', 'fb_cpg.h', 'fb_drag.h')]
    sources.append(os.path.join(ROOT, 'include', 'farms_b200.h'))
    stale = (not os.path.exists(EMU_LIB)
             or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in sources))
    if stale:
        os.makedirs(EMU_DIR, exist_ok=True)
        subprocess.run(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-DFB_HOST_EMU', '-x', 'c++',
                        sources[0], '-o', EMU_LIB], check=True)
    return EMU_LIB


@pytest.fixture(scope='session')
def cuda_library():
    """The product library; GPU tests fail loudly if it is missing."""
    import __graft_entry__ as entry
    return entry.build_engine()


def make_case(name, n_envs, seed=0, qvel_scale=0.3, ctrl_scale=0.3):
    """Seeded inputs shared by the emulation and GPU parity tests."""
    from farms_mujoco_b200 import models, mjcf_subset
    spec = models.MODELS[name]()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    rng = np.random.default_rng(seed)
    qpos0 = np.tile(model.key_qpos, (n_envs, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, size=(n_envs, model.nq - 7))
    qvel0 = np.tile(model.key_qvel, (n_envs, 1)) + rng.uniform(
        -qvel_scale, qvel_scale, size=(n_envs, model.nv))
    ctrl = rng.uniform(-ctrl_scale, ctrl_scale, size=(n_envs, model.nu))
    return spec, model, qpos0, qvel0, ctrl


def oracle_rollout(spec, model, tables, n_rows, qpos0, qvel0, ctrl):
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    physics = OraclePhysics(model)
    data, states = fo.reference_rollout(
        physics, spec, tables, n_rows, controller=lambda it, t: ctrl, qpos0=qpos0, qvel0=qvel0)
    return physics, data, states


def scaled_error(a, b, floor=1.0):
    """max |a - b| / max(max |b|, floor): error relative to the largest magnitude of the array,
    absolute in units of `floor` when the whole array is smaller than that."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max()/max(floor, np.abs(b).max()))


# ---- parity metric of the oracle comparisons ---------------------------------------------
# Every array is split into groups of ONE physical quantity (joint angles, root position, link
# linear velocities, contact forces ...) and each group is measured relative to the largest
# magnitude that quantity takes in the reference array:
#       err(group) = max |ours - ref| / max(max |ref|, PARITY_FLOOR),
# the error of the array being the worst of its groups.  It is a relative error throughout (a
# 0.1 rad joint angle is held to tol * 0.1 rad, not to tol * 1); the absolute floor only keeps
# groups that are identically zero in the reference (a limit force while no limit is active)
# from dividing by zero: there the error is absolute, in units of 1e-2 SI (0.01 rad, 1 cm, 0.01 rad/s, 0.01 N).
PARITY_FLOOR = 1e-2


def _log_groups(kind, n_cols):
    from farms_mujoco_b200.layout import sc
    if kind == 'links':
        return {'com_position': slice(0, 3), 'com_orientation': slice(3, 7), 'urdf_position': slice(7, 10),
                'urdf_orientation': slice(10, 14), 'velocity_lin': slice(14, 17), 'velocity_ang': slice(17, 20)}
    if kind == 'joints':
        named = {'position': sc.joint_position, 'velocity': sc.joint_velocity, 'torque': sc.joint_torque,
                 'limit_force': sc.joint_limit_force}
        groups = {k: slice(c, c + 1) for k, c in named.items() if c < n_cols}
        rest = [c for c in range(n_cols) if c not in named.values()]
        if rest:
            groups['unwritten'] = rest           # columns physics.py:481-524 leaves zero
        return groups
    if kind == 'contacts':
        return {'reaction': slice(0, 3), 'friction': slice(3, 6), 'total': slice(6, 9), 'position': slice(9, 12)}
    if kind == 'xfrc':
        return {'force': slice(0, 3), 'torque': slice(3, 6)}
    raise KeyError(kind)


def log_errors(kind, ours, ref):
    """Group-wise relative errors of one log array [..., n_items, n_cols] (see PARITY_FLOOR)."""
    ours, ref = np.asarray(ours), np.asarray(ref)
    assert ours.shape == ref.shape, (kind, ours.shape, ref.shape)
    return {name: scaled_error(ours[..., cols], ref[..., cols], PARITY_FLOOR)
            for name, cols in _log_groups(kind, ref.shape[-1]).items()}


def log_error(kind, ours, ref):
    return max(log_errors(kind, ours, ref).values(), default=0.0)


def state_errors(qpos, qvel, ref_qpos, ref_qvel):
    """Group-wise relative errors of a state: root position / root quaternion / joint positions and
    root linear / root angular / joint velocities (free base: nq == nv + 1), else one group each."""
    qpos, qvel, ref_qpos, ref_qvel = (np.asarray(x, dtype=np.float64) for x in (qpos, qvel, ref_qpos, ref_qvel))
    if qpos.shape[-1] == qvel.shape[-1] + 1:
        groups = {'root_position': (qpos[..., :3], ref_qpos[..., :3]), 'root_quaternion': (qpos[..., 3:7], ref_qpos[..., 3:7]),
                  'joint_position': (qpos[..., 7:], ref_qpos[..., 7:]), 'root_velocity_lin': (qvel[..., :3], ref_qvel[..., :3]),
                  'root_velocity_ang': (qvel[..., 3:6], ref_qvel[..., 3:6]), 'joint_velocity': (qvel[..., 6:], ref_qvel[..., 6:])}
    else:
        groups = {'joint_position': (qpos, ref_qpos), 'joint_velocity': (qvel, ref_qvel)}
    return {name: scaled_error(a, b, PARITY_FLOOR) for name, (a, b) in groups.items()}


def parity_errors(qpos, qvel, logs, ref_qpos, ref_qvel, ref_data, rows=None):
    """{'qpos': e, 'qvel': e, 'links': e, ...} -- worst group of each array, plus the per-group
    detail under 'detail'.  `logs[kind]` is [n_rows, n_items, n_cols] of ONE environment."""
    st = state_errors(qpos, qvel, ref_qpos, ref_qvel)
    out = {'qpos': max(v for k, v in st.items() if 'velocity' not in k),
           'qvel': max(v for k, v in st.items() if 'velocity' in k)}
    detail = dict(st)
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        ref = getattr(ref_data.sensors, kind).array
        ours = logs[kind]
        if rows is not None:
            ref, ours = ref[rows], ours[rows]
        errs = log_errors(kind, ours, ref)
        out[kind] = max(errs.values(), default=0.0)
        detail.update({f'{kind}.{k}': v for k, v in errs.items()})
    out['detail'] = detail
    return out
