"""Device code (fb_device.h) compiled for the host (TEAM = 1) vs the fp64 oracle.

This checks the fp32 arithmetic, the indexing and the C-ABI marshalling of the
CUDA path on the GPU-less box.  It is NOT the product path and proves nothing
about parallel execution: the `-m gpu` tests do that on a B200.
"""

import numpy as np
import pytest

from conftest import make_case, oracle_rollout, scaled_error, log_error, state_errors


def _run(emu_library, name, n_envs, n_steps, **kw):
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs, **kw)
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=emu_library)
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps, want_derived=True)
    return spec, model, qpos0, qvel0, ctrl, physics


@pytest.mark.parametrize('name,tol', [('swimmer8', 5e-4), ('salamander_swim', 5e-4),
                                      ('salamander', 5e-3), ('centipede', 5e-3)])
def test_rollout_matches_oracle(emu_library, name, tol):
    n_steps = 12
    spec, model, qpos0, qvel0, ctrl, physics = _run(emu_library, name, 3, n_steps)
    logs = physics.log_arrays()
    assert not physics.flags.any()
    for env in range(3):
        _, data, states = oracle_rollout(spec, model, physics.tables, n_steps + 1, qpos0[env],
                                         qvel0[env], ctrl[env])
        errs = state_errors(physics.qpos[env], physics.qvel[env], *states[-1])
        assert max(errs.values()) < tol, errs
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert log_error(kind, logs[kind][env], getattr(data.sensors, kind).array) < tol, kind


@pytest.mark.parametrize('name,tol', [('swimmer8', 2e-5), ('salamander_swim', 2e-5),
                                      ('salamander', 2e-5), ('centipede', 2e-5)])
def test_fast_path_rollout_matches_oracle(emu_library, name, tol):
    """Default fb_step: environment-per-thread kernel, then the per-thread constrained kernel
    (matrix-free Newton, fb_fastc.h) on the environments with an active limit / contact."""
    import fastpath_cases
    from farms_mujoco_b200.engine import BatchedPhysics
    n_steps = 12
    spec, model, qpos0, qvel0, ctrl = make_case(name, 3)
    physics = BatchedPhysics.from_spec(spec, 3, buffer_size=n_steps + 1, library=emu_library)
    assert physics.fast_path and physics.constraint_path == 1
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.last_pending == (3 if name in ('salamander', 'centipede') else 0)
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(3), n_steps, tol,
                                       tol_contacts=5e-4)


@pytest.mark.parametrize('name', ['salamander', 'centipede'])
def test_constrained_single_step_within_1e_5(emu_library, name):
    """BASELINE.json north_star: single-step qpos / qvel within 1e-5.  Contact forces: 2e-4 --
    they are proportional to the penetration depth (1e-5 .. 1e-3 m), a difference of O(0.1 m)
    positions that the fp32 state itself only holds to 4e-9 m."""
    import fastpath_cases
    from farms_mujoco_b200.engine import BatchedPhysics
    n = 12
    spec, model, qpos0, qvel0, ctrl = make_case(name, n)
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=2, library=emu_library)
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(1)
    assert physics.last_pending == n
    assert physics.log_arrays()['contacts'].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(n), 1, 1e-5,
                                       tol_contacts=2e-4)


def test_team_kernel_on_hand_overs(emu_library):
    """fb_set_constraint_path(0): the team kernel (CRB + L'DL + Newton) finishes the hand-overs."""
    import fastpath_cases
    from farms_mujoco_b200.engine import BatchedPhysics
    n_steps = 12
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', 3)
    physics = BatchedPhysics.from_spec(spec, 3, buffer_size=n_steps + 1, library=emu_library)
    physics.set_constraint_path(False)
    assert physics.constraint_path == 0
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.last_pending == 3
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(3), n_steps, 5e-3)


@pytest.mark.parametrize('per_thread', [True, False])
def test_fast_path_hand_over(emu_library, per_thread):
    import fastpath_cases
    fastpath_cases.check_hand_over(emu_library, 'swimmer8', n_envs=9, per_thread=per_thread)


def test_constraint_paths_agree(emu_library):
    """Ground contact: per-thread constrained kernel vs team kernel on the same rollout."""
    import fastpath_cases
    fastpath_cases.check_constraint_paths_agree(emu_library, 'salamander', n_envs=3)


def test_fast_and_team_paths_agree(emu_library):
    import fastpath_cases
    fastpath_cases.check_paths_agree(emu_library, 'salamander_swim', n_envs=2)


def test_fast_path_joint_and_actuator_variants(emu_library):
    """Slide joint, off-origin anchors, tilted axes, stiffness + springref, rotated body
    frame, general inertia, ctrl / force clamps, geared motor: 15 steps within 2e-5."""
    import fastpath_cases
    import variant_models
    fastpath_cases.check_variant(emu_library, variant_models.swimmer8_features())


def test_fast_path_fixed_base(emu_library):
    import fastpath_cases
    import variant_models
    fastpath_cases.check_variant(emu_library, variant_models.swimmer8_fixed_base(), free_base=False)


@pytest.mark.parametrize('which', ['swimmer8', 'features', 'salamander'])
def test_ctrl_sequence(emu_library, which):
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import models
    spec = {'swimmer8': models.swimmer8, 'features': variant_models.swimmer8_features,
            'salamander': models.salamander}[which]()
    fastpath_cases.check_ctrl_sequence(emu_library, spec)


def test_log_layout_is_reference_layout(emu_library):
    """Row k: links/contacts of state k-1 (k=0: state 0), joints qpos/qvel of state k
    (SURVEY.md Appendix D-1); quaternions xyzw; unwritten joint columns stay zero."""
    from farms_mujoco_b200.layout import sc
    spec, model, qpos0, qvel0, ctrl, physics = _run(emu_library, 'salamander_swim', 2, 3)
    logs = physics.log_arrays(env=1)
    links, joints = logs['links'], logs['joints']
    assert links.shape == (4, 28, 20) and joints.shape == (4, 27, 18)
    assert np.array_equal(links[0], links[1])                      # both show state 0
    assert not np.array_equal(links[1], links[2])
    assert np.allclose(joints[0, :, sc.joint_position], qpos0[1, 7:], atol=1e-7)
    assert np.allclose(joints[3, :, sc.joint_position], physics.qpos[1, 7:], atol=0)
    written = [sc.joint_position, sc.joint_velocity, sc.joint_torque, sc.joint_limit_force]
    others = [c for c in range(sc.joint_size) if c not in written]
    assert not joints[:, :, others].any()
    assert np.allclose(np.linalg.norm(links[:, :, 3:7], axis=-1), 1, atol=1e-6)
    assert np.array_equal(links[:, :, 3:7], links[:, :, 10:14])    # D-3
    exported = physics.export_farms(1)
    assert exported.sensors.links.array.dtype == np.float64
    assert np.array_equal(exported.sensors.links.array, links.astype(np.float64))


def test_units_scaling(emu_library):
    """MJCF authored in scaled units, logs in SI (physics.py:428-523): a model
    scaled by `meters` logs the same SI numbers."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.units import SimulationUnitScaling
    spec = models.swimmer8()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    base = BatchedPhysics.from_spec(spec, 1, buffer_size=2, library=emu_library)
    units = SimulationUnitScaling(meters=1.0, seconds=1.0, kilograms=1.0)
    spec.simulation_options.units = units
    same = BatchedPhysics.from_spec(spec, 1, buffer_size=2, library=emu_library)
    assert np.array_equal(base.log_arrays()['links'], same.log_arrays()['links'])
    assert base.tables.meters == 1.0 and model.nv == 13


def test_wave_controller_equals_host_controller(emu_library):
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.models import travelling_wave_parameters
    spec, model, qpos0, qvel0, _ = make_case('swimmer8', 2)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    phase = np.array([0.3, 1.7])
    dev = BatchedPhysics.from_spec(spec, 2, buffer_size=6, library=emu_library)
    dev.set_env_phase(phase)
    dev.set_wave_controller(acts, amp, freq, lag)
    dev.reset(qpos0, qvel0)
    dev.step(5)
    host = BatchedPhysics.from_spec(spec, 2, buffer_size=6, library=emu_library)
    host.reset(qpos0, qvel0)
    for it in range(5):
        ctrl = np.zeros((2, model.nu))
        ctrl[:, acts] = amp*np.sin(2*np.pi*freq*it*model.timestep - lag + phase[:, None])
        host.set_ctrl(ctrl)
        host.step(1)
    assert np.allclose(dev.qpos, host.qpos, atol=2e-6)
    assert np.allclose(dev.qvel, host.qvel, atol=2e-4)


def test_swimming_switches(emu_library):
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('swimmer8', 1)
    physics = BatchedPhysics.from_spec(spec, 1, buffer_size=3, library=emu_library)
    physics.set_swimming(False, False)
    physics.reset(qpos0, qvel0)
    physics.step(2)
    assert not physics.log_arrays()['xfrc'].any() and not physics.xfrc_applied.any()
    physics.set_swimming(True, True)
    physics.set_water_velocity([0.1, 0.0, 0.0])
    physics.reset(qpos0, qvel0)
    physics.step(2)
    assert physics.log_arrays()['xfrc'].any() and physics.xfrc_applied.any()


def test_errors_are_reported(emu_library):
    from farms_mujoco_b200.engine import BatchedPhysics, EngineError
    spec, model, qpos0, qvel0, ctrl = make_case('swimmer8', 1)
    physics = BatchedPhysics.from_spec(spec, 1, library=emu_library)
    with pytest.raises(EngineError, match='n_steps'):
        physics.step(0)
    with pytest.raises(EngineError, match='actuator'):
        physics.set_wave_controller([999], [1.0], [1.0], [0.0])


def test_step_host_joint_columns(emu_library):
    """fb_set_host_joint_columns: the joints row comes down as the selected columns."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.layout import sc
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', 3)
    nl, nj = len(spec.links_names), len(spec.joints_names)
    cols = [sc.joint_position, sc.joint_velocity, sc.joint_torque, sc.joint_limit_force]
    rows = {}
    for compact in (False, True):
        physics = BatchedPhysics.from_spec(spec, 3, buffer_size=8, library=emu_library)
        physics.reset(qpos0, qvel0)
        links = np.zeros((3, nl, 20), dtype=np.float32)
        joints = np.zeros((3, nj, len(cols) if compact else sc.joint_size), dtype=np.float32)
        if compact:
            physics.set_host_joint_columns(cols)
        physics.step_host(4, ctrl=ctrl.astype(np.float32), links_row=links, joints_row=joints)
        rows[compact] = (links, joints)
    assert np.array_equal(rows[True][0], rows[False][0])
    assert np.array_equal(rows[True][1], rows[False][1][:, :, cols])
    assert rows[True][1][:, :, :3].any()


@pytest.mark.parametrize('name', ['swimmer8', 'salamander_swim', 'salamander'])
def test_slim_layout_is_bit_identical(emu_library, name):
    """The large-batch (SLIM) layout of the unconstrained kernel keeps velocities and slots in the
    scratch instead of shared memory: same arithmetic, bit-identical results (hand-overs too)."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, 3)
    outs = []
    for slim in (0, 1, 8):
        physics = BatchedPhysics.from_spec(spec, 3, buffer_size=9, library=emu_library)
        physics.set_fast_slim(slim)
        assert physics.fast_slim == slim
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(5)
        physics.step(3)
        outs.append((physics.qpos, physics.qvel, physics.xfrc_applied, physics.log_arrays()))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and np.array_equal(outs[0][1], other[1])
        assert np.array_equal(outs[0][2], other[2])
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.array_equal(outs[0][3][kind], other[3][kind]), kind


@pytest.mark.parametrize('per_thread,tol', [(True, 2e-5), (False, 5e-4)])
def test_box_plane_contacts(emu_library, per_thread, tol):
    """Plane-box contacts (mjc_PlaneBox corners): SALAMANDER with box feet and a
    box trunk segment, both constraint paths vs the oracle."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_box_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert set(range(2, 10)) <= set(model.cand_end.tolist())
    n, n_steps = 4, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=emu_library)
    physics.set_constraint_path(per_thread)
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.last_pending == n and physics.log_arrays()['contacts'].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(n), n_steps, tol,
                                       tol_contacts=max(tol, 2e-4))


@pytest.mark.parametrize('which', ['features', 'fixed_base'])
def test_slim_layout_variants_and_ctrl_sequence(emu_library, which):
    """SLIM layout on the hand-edited models (slide joint, off-origin anchors, clamps, fixed base)
    with an uploaded control sequence: bit-identical to the regular layout."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics

    spec = variant_models.swimmer8_features() if which == 'features' else variant_models.swimmer8_fixed_base()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n, n_steps = 3, 7
    rng = np.random.default_rng(2)
    first = 7 if which == 'features' else 0
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, first:] += rng.uniform(-0.1, 0.1, (n, model.nq - first))
    qvel0 = rng.uniform(-0.3, 0.3, (n, model.nv))
    seq = rng.uniform(-0.3, 0.3, (n_steps, n, model.nu)).astype(np.float32)
    outs = []
    for slim in (0, 1, 8):
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=emu_library)
        physics.set_fast_slim(slim)
        physics.reset(qpos0, qvel0)
        physics.set_ctrl_sequence(seq)
        physics.step(3)
        physics.step(n_steps - 3)
        outs.append((physics.qpos, physics.qvel, physics.ctrl, physics.log_arrays()))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and np.array_equal(outs[0][1], other[1])
        assert np.array_equal(outs[0][2], other[2])
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.array_equal(outs[0][3][kind], other[3][kind]), kind


@pytest.mark.parametrize('kind', ['limits', 'contacts'])
def test_ring_wrap(emu_library, kind):
    """buffer_size < n_steps: the ring wraps five times."""
    import fastpath_cases
    fastpath_cases.check_ring_wrap(emu_library, kind, n_envs=4)


def test_reset_clears_constraint_columns(emu_library):
    import fastpath_cases
    fastpath_cases.check_reset_clears_log(emu_library)


def test_host_link_columns_and_row_export(emu_library):
    """fb_set_host_link_columns (links row as selected columns) and fb_export_rows (whole ring
    rows of every environment, dense) against the per-environment log view."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.layout import sc
    n, n_steps, ring = 5, 7, 4
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', n)
    nl, nj = len(spec.links_names), len(spec.joints_names)
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=ring, library=emu_library)
    physics.reset(qpos0, qvel0)
    cols = list(range(sc.link_com_position_x, sc.link_com_orientation_w + 1))
    physics.set_host_link_columns(cols)
    links = np.zeros((n, nl, len(cols)), np.float32)
    joints = np.zeros((n, nj, sc.joint_size), np.float32)
    physics.step_host(n_steps, ctrl=np.ascontiguousarray(ctrl, dtype=np.float32), links_row=links, joints_row=joints)
    logs = physics.log_arrays()
    last = n_steps % ring
    assert np.array_equal(links, logs['links'][:, last][:, :, cols])
    assert np.array_equal(joints, logs['joints'][:, last])
    physics.set_host_link_items([0, 5])
    two = np.zeros((n, 2, len(cols)), np.float32)
    # ... and ctrl goes up as three selected actuators, the others keeping their value
    acts = [2, 7, 11]
    physics.set_host_ctrl_columns(acts)
    some = np.ascontiguousarray(np.arange(n*3, dtype=np.float32).reshape(n, 3)*1e-3)
    physics.step_host(1, ctrl=some, links_row=two)
    want_ctrl = np.array(ctrl, dtype=np.float32)
    want_ctrl[:, acts] = some
    assert np.array_equal(physics.ctrl, want_ctrl)
    physics.set_host_ctrl_columns(None)
    logs = physics.log_arrays()
    assert np.array_equal(two, logs['links'][:, (n_steps + 1) % ring][:, [0, 5]][:, :, cols])
    physics.set_host_link_items(None)
    physics.set_host_link_columns(None)
    n_steps += 1
    full = np.zeros((n, nl, 20), np.float32)
    physics.step_host(1, links_row=full)
    logs = physics.log_arrays()
    assert np.array_equal(full, logs['links'][:, (n_steps + 1) % ring])
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        out = physics.export_rows(kind, 2, 4)           # rows 2, 3, 0, 1
        want = logs[kind][:, [2, 3, 0, 1]].transpose(1, 0, 2, 3)
        assert out.shape == want.shape and np.array_equal(out, want), kind
    assert logs['contacts'].any()


def test_device_cpg_matches_host_controller(emu_library):
    """coupled-oscillator CPG + torque writer on the device == step_control on the host."""
    import fastpath_cases
    fastpath_cases.check_device_cpg(emu_library)


def test_lean_variant(emu_library):
    import fastpath_cases
    fastpath_cases.check_lean_variant(emu_library, slims=(0, 1))


def test_split_variant_is_bit_identical(emu_library):
    """SPLIT variant of the unconstrained kernel: the warps of a block emulated by host threads that
    meet at the kernel's barriers (one lane each) -- the schedule, the phase barriers and the
    hand-over protocol are the device's."""
    import fastpath_cases
    fastpath_cases.check_split_variant_is_bit_identical(emu_library, n_envs=9)


def test_con_split_variant(emu_library):
    """SPLIT variant of the constrained kernel under the same emulation (cross-warp sums included)."""
    import fastpath_cases
    fastpath_cases.check_con_split_variant(emu_library, n_envs=75, n_steps=(3, 2))


def test_split_variant_fixed_base(emu_library):
    """Tree-split kernel (emulated warps) on a branching tree without a floating root."""
    import fastpath_cases
    import variant_models
    fastpath_cases.check_variant(emu_library, variant_models.salamander_swim_fixed_base(), free_base=False)


def test_con_split_mixed_groups(emu_library):
    import fastpath_cases
    fastpath_cases.check_con_split_mixed_groups(emu_library, n_envs=16)


@pytest.mark.parametrize('per_thread,tol', [(True, 2e-5), (False, 5e-4)])
def test_cylinder_plane_contacts(emu_library, per_thread, tol):
    """Plane-cylinder contacts (mjc_PlaneCylinder: up to four points per cylinder):
    SALAMANDER with tilted cylinder feet and a cylinder trunk segment, both constraint paths (and the
    SPLIT constrained kernel) vs the oracle."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_cylinder_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert sum(e >= 11 for e in model.cand_end.tolist()) == 20          # four feet and a trunk segment x 4 points
    n, n_steps = 4, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    for split in ((False, True) if per_thread else (False,)):
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=emu_library)
        physics.set_constraint_path(per_thread)
        physics.set_con_split(split)
        assert not physics.fast_lean
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(n_steps)
        logs = physics.log_arrays()
        assert physics.last_pending == n and logs['contacts'][:, :, 12:].any()
        fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(n), n_steps, tol,
                                           tol_contacts=max(tol, 2e-4))


@pytest.mark.parametrize('per_thread,tol', [(True, 2e-5), (False, 5e-4)])
def test_ellipsoid_plane_contacts(emu_library, per_thread, tol):
    """Plane-ellipsoid contacts (mjc_PlaneConvex: one contact at the ellipsoid's support point along
    the plane normal): SALAMANDER with rotated ellipsoid feet and an ellipsoid trunk
    segment, both constraint paths (and the SPLIT constrained kernel) vs the oracle."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_ellipsoid_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert model.cand_end.tolist().count(10) == 5          # four feet and a trunk segment
    n, n_steps = 4, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    for split in ((False, True) if per_thread else (False,)):
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=emu_library)
        physics.set_constraint_path(per_thread)
        physics.set_con_split(split)
        assert not physics.fast_lean          # the LEAN variants leave the ellipsoid branch out
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(n_steps)
        logs = physics.log_arrays()
        assert physics.last_pending == n and logs['contacts'][:, :, 12:].any()      # the feet's sensors
        fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(n), n_steps, tol,
                                           tol_contacts=max(tol, 2e-4))


def test_drag_operator(emu_library):
    """fb_drag_forces (the stand-alone drag_forces of drag.pyx:152-268, float64) against the oracle:
    1e-12 relative (same formulae, other order of the quaternion products)."""
    from drag_cases import check_operator
    assert check_operator(emu_library) < 1e-12


def test_drag_forces_reference_signature(emu_library, monkeypatch):
    """swimming.drag.drag_forces with the reference's argument list on an AnimatData row."""
    from farms_mujoco_b200 import engine
    from farms_mujoco_b200.swimming.drag import drag_forces, WaterProperties
    from drag_cases import make_rows, oracle_answer
    monkeypatch.setattr(engine, 'DEFAULT_LIBRARY', emu_library)
    case = make_rows(6)

    class _Arr:      # the two attributes drag_forces touches
        def __init__(self, array):
            self.array = array
    links, xfrc = _Arr(case['links'][None].copy()), _Arr(np.full((1, 6, 6), 0.25))
    water = WaterProperties(case['surface'], 1000.0, case['wvel'], case['viscosity'])
    ref, ref_applied = oracle_answer(case, True, xfrc.array[0])
    for i in range(6):
        got = drag_forces(0, links, i, xfrc, i, case['coef'][i], None, None, water, case['mass'][i],
                          case['height'][i], case['density'][i], case['gravity'], True)
        assert got == ref_applied[i]
    assert np.abs(xfrc.array[0] - ref).max() < 1e-12


def _pair_case(n, seed=5):
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    spec = variant_models.salamander_foot_pairs()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    rng = np.random.default_rng(seed)
    qpos0 = np.tile(variant_models.folded_legs_qpos(model, 1.0), (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.03, 0.03, (n, model.nq - 7))
    qpos0[n - 1] = model.key_qpos                      # one environment with the legs apart: no pair contact
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = np.tile(qpos0[0, 7:][None], (n, 1))[:, :0]  # placeholder, replaced below
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    return spec, model, qpos0, qvel0, ctrl


def test_pair_self_collisions(emu_library):
    """Explicit <contact><pair> self-collisions (mjcf.py:1012-1033; sphere-sphere, mjc_SphereSphere):
    rows on two branches of the tree -> the team kernel with the dense Newton Hessian, vs the
    oracle.  The pair sensor and the two single-link sensors see the contact with opposite signs."""
    import fastpath_cases
    from farms_mujoco_b200.engine import BatchedPhysics
    n, n_steps = 3, 12
    spec, model, qpos0, qvel0, ctrl = _pair_case(n)
    assert (model.cand_end == 20).sum() == 2
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=emu_library)
    assert physics.fast_path == 0                       # two-body rows: team kernel only
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    contacts = physics.log_arrays()['contacts']
    names = [tuple(c) for c in spec.contacts_names]
    pair = names.index(('link_leg_0_L_3', 'link_leg_0_R_3'))
    left, right = names.index(('link_leg_0_L_3', '')), names.index(('link_leg_0_R_3', ''))
    assert np.abs(contacts[0, 1:, pair, 6:9]).max() > 1e-3          # the feet push on one another
    assert not contacts[n - 1, :, pair].any()                        # legs apart: nothing
    # (g1, g2): -1 and (g1, -1): -1 on the left link, (g2, -1): +1 on the right (sensors.pyx:160-176)
    assert np.allclose(contacts[0, :, pair, 6:9], contacts[0, :, left, 6:9], atol=1e-6)
    assert np.allclose(contacts[0, :, pair, 6:9], -contacts[0, :, right, 6:9], atol=1e-6)
    # team-kernel tolerances (the CRB + L'DL path accumulates more rounding than the recursions)
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, range(n), n_steps, 2e-4,
                                       tol_contacts=2e-4)


def test_pair_friction_floor():
    """The reference's generated pairs are frictionless (friction=[0]*5 -> MuJoCo's floor 1e-5,
    mjcf.py:1029): the pyramidal rows' D = 1/(2 mu^2 R) puts them out of reach of the fp32 solver;
    the front-end says so instead of returning forces that are wrong."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    with pytest.raises(NotImplementedError, match='friction'):
        mjcf_subset.parse_mjcf(variant_models.salamander_foot_pairs(friction=0).mjcf)
    with pytest.raises(NotImplementedError, match='sphere-sphere'):
        spec = variant_models.salamander_foot_pairs()
        mjcf_subset.parse_mjcf(spec.mjcf.replace('geom1="link_leg_0_L_3_foot"', 'geom1="link_body_3_collision"'))


def test_drag_operator_edge_cases(emu_library):
    """fb_drag_forces: no rows is a no-op, a null argument is an error with a message, the Python
    wrapper refuses output arrays it could not update in place."""
    import ctypes as ct
    from farms_mujoco_b200 import cabi
    from farms_mujoco_b200.engine import load_library, drag_forces_rows
    lib = load_library(emu_library)
    null = ct.cast(None, cabi.c_double_p)
    assert lib.fb_drag_forces(0, 0, null, null, null, null, null, 0.0, null, 1.0, -9.81, 1, null, None) == 0
    assert lib.fb_drag_forces(0, 2, null, null, null, null, null, 0.0, null, 1.0, -9.81, 1, null, None) != 0
    assert b'null' in lib.fb_last_error()
    links = np.zeros((2, 20)); links[:, 6] = links[:, 13] = 1.0
    with pytest.raises(ValueError):
        drag_forces_rows(links, np.zeros((2, 6)), 1.0, 0.1, 1000.0, 0.0, [0, 0, 0], 1.0, -9.81, True,
                         np.zeros((2, 6), dtype=np.float32), library=emu_library)
    xfrc = np.zeros((2, 6))
    applied = drag_forces_rows(links, np.zeros((2, 6)), 1.0, 0.1, 1000.0, 0.0, [0, 0, 0], 1.0, -9.81, False,
                               xfrc, library=emu_library)
    assert applied.all() and not xfrc.any()                 # at the surface, at rest, no buoyancy: zero forces
