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
    sources = [os.path.join(CSRC, f) for f in ('fb_engine.cu', 'fb_device.h', 'fb_fast.h', 'fb_fastc.h', 'fb_model.h')]
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


def scaled_error(a, b):
    """max |a-b| / max(1, max|b|): absolute below 1, relative above."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max()/max(1.0, np.abs(b).max()))
