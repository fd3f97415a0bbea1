"""Opportunistic TRUE parity tier (SURVEY.md 8c pin 4): when MuJoCo itself is importable, the
oracle's restatement of the MuJoCo subset (oracle/mjstep_oracle.c) is compared with ``mj_step`` on
the same MJCF and the same state -- the call the reference makes at farms_mujoco/simulation/
simulation.py:53 (model) and :156 (step), and ``mj_contactForce`` of sensors/sensors.pyx:70.

MuJoCo is absent from this image and from /root/reference, so here (and on the GPU boxes, which
run the same image) these tests are reported as SKIPPED; the oracle's header keeps saying
"parity unpinned" until they have run green somewhere.  They need nothing but ``pip install
mujoco`` and this repository.
"""

import numpy as np
import pytest

mujoco = pytest.importorskip('mujoco', reason='MuJoCo is not installed: the true-parity tier is skipped')

from conftest import make_case  # noqa: E402  pylint: disable=wrong-import-position


MODELS = ['swimmer8', 'salamander_swim', 'salamander', 'centipede']


def _both(name, n_steps, seed=0):
    from oracle.oracle import OraclePhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, 1, seed=seed)
    mjm = mujoco.MjModel.from_xml_string(spec.mjcf)
    mjd = mujoco.MjData(mjm)
    assert (mjm.nq, mjm.nv, mjm.nu, mjm.nbody) == (model.nq, model.nv, model.nu, model.nbody)
    mujoco.mj_resetData(mjm, mjd)
    mjd.qpos[:] = qpos0[0]
    mjd.qvel[:] = qvel0[0]
    mjd.ctrl[:] = ctrl[0]
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    orc.data.qpos[:] = qpos0[0]
    orc.data.qvel[:] = qvel0[0]
    orc.data.ctrl[:] = ctrl[0]
    for _ in range(n_steps):
        mujoco.mj_step(mjm, mjd)
        orc.step()
    return model, mjm, mjd, orc


@pytest.mark.parametrize('name', MODELS)
def test_model_constants_match_the_compiler(name):
    """mjcf_subset.parse_mjcf vs MuJoCo's own compiler on the fields the step reads."""
    from farms_mujoco_b200 import mjcf_subset
    spec, model, *_ = make_case(name, 1)
    mjm = mujoco.MjModel.from_xml_string(spec.mjcf)
    for ours, theirs in (('body_mass', mjm.body_mass), ('body_inertia', mjm.body_inertia),
                         ('body_pos', mjm.body_pos), ('body_quat', mjm.body_quat), ('body_ipos', mjm.body_ipos),
                         ('body_iquat', mjm.body_iquat), ('jnt_axis', mjm.jnt_axis), ('jnt_pos', mjm.jnt_pos),
                         ('jnt_range', mjm.jnt_range), ('dof_damping', mjm.dof_damping),
                         ('dof_armature', mjm.dof_armature), ('dof_invweight0', mjm.dof_invweight0),
                         ('body_invweight0', mjm.body_invweight0), ('qpos0', mjm.qpos0)):
        a = np.asarray(getattr(model, ours), dtype=float).reshape(np.asarray(theirs).shape)
        assert np.allclose(a, theirs, rtol=1e-9, atol=1e-12), ours
    assert model.timestep == mjm.opt.timestep
    del mjcf_subset


@pytest.mark.parametrize('name', MODELS)
def test_single_step_matches_mj_step(name):
    model, mjm, mjd, orc = _both(name, 1)
    assert np.allclose(orc.data.qpos, mjd.qpos, rtol=0, atol=1e-10)
    assert np.allclose(orc.data.qvel, mjd.qvel, rtol=1e-8, atol=1e-9)
    # derived quantities of the PRE-step state stay in mjData after mj_step
    assert np.allclose(orc.data.xpos, mjd.xpos, atol=1e-12)
    assert np.allclose(orc.data.xquat, mjd.xquat, atol=1e-12)
    assert np.allclose(orc.data.xipos, mjd.xipos, atol=1e-12)
    assert np.allclose(orc.data.actuator_force, mjd.actuator_force, rtol=1e-9, atol=1e-12)
    assert orc.ncon == mjd.ncon
    for i in range(mjd.ncon):
        force = np.zeros(6)
        mujoco.mj_contactForce(mjm, mjd, i, force)
        # contact order can differ between the two collision passes: match by geometry pair
        pair = (int(mjd.contact[i].geom1), int(mjd.contact[i].geom2))
        ours = [k for k, c in enumerate(orc.data.contact) if (c.geom1, c.geom2) == pair
                and np.allclose(c.pos, mjd.contact[i].pos, atol=1e-9)]
        assert len(ours) == 1, pair
        assert np.allclose(orc.contact_force(ours[0])[:3], force[:3], rtol=1e-6, atol=1e-9)
    del model


@pytest.mark.parametrize('name', MODELS)
def test_hundred_steps_match_mj_step(name):
    _, _, mjd, orc = _both(name, 100, seed=1)
    assert np.allclose(orc.data.qpos, mjd.qpos, rtol=0, atol=1e-7)
    assert np.allclose(orc.data.qvel, mjd.qvel, rtol=1e-6, atol=1e-6)


def test_pair_self_collision_matches_mj_step():
    """Explicit <contact><pair> sphere-sphere contact (mjcf.py:1012-1033 with a friction coefficient):
    contact geometry, the two-body rows (mj_jacDifPair) and mj_contactForce against MuJoCo."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from oracle.oracle import OraclePhysics
    spec = variant_models.salamander_foot_pairs()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    mjm = mujoco.MjModel.from_xml_string(spec.mjcf)
    mjd = mujoco.MjData(mjm)
    qpos = variant_models.folded_legs_qpos(model, 1.0)
    mujoco.mj_resetData(mjm, mjd)
    mjd.qpos[:] = qpos
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    orc.data.qpos[:] = qpos
    mujoco.mj_step(mjm, mjd)
    orc.step()
    assert orc.ncon == mjd.ncon == 2
    assert np.allclose(orc.data.qvel, mjd.qvel, rtol=1e-8, atol=1e-9)
    for i in range(mjd.ncon):
        force = np.zeros(6)
        mujoco.mj_contactForce(mjm, mjd, i, force)
        ours = [k for k, c in enumerate(orc.data.contact)
                if {c.geom1, c.geom2} == {int(mjd.contact[i].geom1), int(mjd.contact[i].geom2)}]
        assert len(ours) == 1
        c = orc.data.contact[ours[0]]
        assert np.allclose(c.pos, mjd.contact[i].pos, atol=1e-12) and abs(c.dist - mjd.contact[i].dist) < 1e-12
        assert np.allclose(c.frame[:3], mjd.contact[i].frame[:3], atol=1e-12)
        assert np.allclose(orc.contact_force(ours[0])[:3], force[:3], rtol=1e-6, atol=1e-9)
