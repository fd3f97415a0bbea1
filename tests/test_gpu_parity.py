"""CUDA path vs the fp64 oracle, through the C ABI, on the same seeded inputs.

Three kernel selections are checked (include/farms_b200.h, fb_set_fast_path /
fb_set_constraint_path):
  * 'fast': the default fb_step -- the environment-per-thread kernel (articulated-body
    recursion) plus the per-thread constrained kernel (matrix-free Newton, fb_fastc.h) on
    whatever it hands over.  Single step 1e-5 (BASELINE.json north_star) on every model, contact
    forces included for the SALAMANDER (CENTIPEDE: 1e-4, see CONTACT_TOL); 20 steps 5e-5.
  * 'fast_team': the per-thread kernel plus the TEAM kernel on the hand-overs.
  * 'team': the team kernel alone (CRB + L'DL + constraint solver, as MuJoCo does it), a
    secondary path: TEAM_TOL.
Metric: group-wise relative error, conftest.py (PARITY_FLOOR).  Measured values in DESIGN.md 5.
Kernel variants (layouts, LEAN, SPLIT, block sizes) are held to each other bit for bit where they
share a compilation and to a few ulp where they do not.
"""

import numpy as np
import pytest

from conftest import make_case, oracle_rollout, scaled_error, parity_errors, log_error

pytestmark = pytest.mark.gpu

MODELS = ['swimmer8', 'salamander_swim', 'salamander', 'centipede']


SWIMMING = ('swimmer8', 'salamander_swim')


def _run(cuda_library, name, n_envs, n_steps, team=0, seed=0, path='team', check_pending=True, **kw):
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs, seed=seed, **kw)
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, team_lanes=team,
                                       library=cuda_library)
    physics.set_fast_path(path != 'team')
    physics.set_constraint_path(path == 'fast')
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    if path != 'team':
        assert physics.fast_path in (16, 32) and physics.constraint_path == (path == 'fast')
        physics.step(n_steps)
        # swimming models stay unconstrained here; ground models are handed over at step 0
        if check_pending:
            assert physics.last_pending == (0 if name in SWIMMING else n_envs)
    else:
        assert physics.fast_path == 0
        physics.step(n_steps, want_derived=True)
    return spec, model, qpos0, qvel0, ctrl, physics


def _compare(spec, model, physics, qpos0, qvel0, ctrl, envs, n_steps, tol, tol_contacts=None):
    qpos, qvel, logs = physics.qpos, physics.qvel, physics.log_arrays()
    assert not physics.flags.any(), physics.flags
    worst, detail = {}, {}
    for env in envs:
        _, data, states = oracle_rollout(spec, model, physics.tables, n_steps + 1, qpos0[env],
                                         qvel0[env], ctrl[env])
        ref_q, ref_v = states[-1]
        errs = parity_errors(qpos[env], qvel[env], {k: v[env] for k, v in logs.items()}, ref_q, ref_v, data)
        for key, val in errs.pop('detail').items():
            detail[key] = max(detail.get(key, 0.0), val)
        for key, val in errs.items():
            worst[key] = max(worst.get(key, 0.0), val)
    print(spec.name, n_steps, 'steps:', {k: f'{v:.2e}' for k, v in worst.items()},
          'worst group:', max(detail, key=detail.get))
    for key, val in worst.items():
        assert val < (tol_contacts if key == 'contacts' and tol_contacts else tol), (key, val, worst, detail)


# Team-kernel tolerances (CRB + L'DL in fp32 on mass matrices of condition 3e4 .. 1e6), group-wise
# relative metric of conftest.py: the small groups (root angular velocity next to 50 rad/s limb
# velocities) carry the error.  Measured r2a: swimmer8 1.3e-4, salamander 3.6e-4, centipede 2.2e-3.
TEAM_TOL = {'swimmer8': 5e-4, 'salamander_swim': 1e-3, 'salamander': 1e-3, 'centipede': 1e-2}
# Contact forces on the default (per-thread) path, single step: the north_star's 1e-5 holds for the
# SALAMANDER (measured 1.9e-6 .. 8.6e-6); the CENTIPEDE's 84 candidates reach 3e-5 .. 7e-5 -- a force
# is proportional to a penetration depth of 1e-5 .. 1e-3 m, itself a difference of O(0.1 m) fp32
# positions (ulp 7e-9 m) -- and are held to 1e-4.
CONTACT_TOL = {'salamander': 1e-5, 'centipede': 1e-4}


@pytest.mark.parametrize('path', ['fast', 'fast_team', 'team'])
@pytest.mark.parametrize('name', MODELS)
def test_single_step(cuda_library, name, path):
    n_envs = 70
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, name, n_envs, 1, path=path)
    per_thread = path == 'fast' or (path == 'fast_team' and name in SWIMMING)
    tol = 1e-5 if per_thread else TEAM_TOL[name]
    if path == 'fast' and name not in SWIMMING:
        assert physics.log_arrays()['contacts'].any()
    _compare(spec, model, physics, qpos0, qvel0, ctrl, [0, 1, 17, 65, n_envs - 1], 1, tol,
             tol_contacts=CONTACT_TOL.get(name, tol) if per_thread else max(tol, 5e-3))


@pytest.mark.parametrize('path', ['fast', 'fast_team', 'team'])
@pytest.mark.parametrize('name,tol', [('swimmer8', 2e-3), ('salamander_swim', 2e-3),
                                      ('salamander', 2e-2), ('centipede', 2e-2)])
def test_twenty_steps(cuda_library, name, tol, path):
    n_envs = 33
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, name, n_envs, 20, path=path)
    if path == 'fast' or (path == 'fast_team' and name in SWIMMING):
        tol = 5e-5
    _compare(spec, model, physics, qpos0, qvel0, ctrl, [0, 16, n_envs - 1], 20, tol,
             tol_contacts=max(tol, 5e-4) if tol == 5e-5 else 5e-2)


# 100-step horizon (BASELINE.json north_star: "short-horizon (100-step) trajectories must agree
# within a stated tolerance"), default kernels.  Stated tolerance, scaled error as above:
# swimming 2e-5 for every quantity; ground contact 1e-4 for state and links / joints rows and
# 5e-4 for the contact-force rows.  Measured: swimming <= 3.2e-6, SALAMANDER
# on ground <= 3.6e-6, CENTIPEDE 2.2e-5 (qvel) / 7.3e-5 (contact forces).
HUNDRED = {'swimmer8': (2e-5, 2e-5), 'salamander_swim': (2e-5, 2e-5),
           'salamander': (1e-4, 5e-4), 'centipede': (1e-4, 5e-4)}


@pytest.mark.parametrize('name', MODELS)
def test_hundred_steps(cuda_library, name):
    n_envs = 40
    tol, tol_contacts = HUNDRED[name]
    # (with a constant ctrl some swimmers reach a joint limit within 100 steps and are handed over)
    # ctrl_scale 0.1: at the 0.3 of the shorter tests a few environments leave the stable range of
    # MuJoCo's explicit actuator feedback within 100 steps, in the fp64 oracle too (next test)
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, name, n_envs, 100, path='fast',
                                                    check_pending=False, ctrl_scale=0.1)
    _compare(spec, model, physics, qpos0, qvel0, ctrl, [0, 13, n_envs - 1], 100, tol,
             tol_contacts=tol_contacts)


def test_divergence_flags_agree_with_the_oracle(cuda_library):
    """Random ctrl of +-0.3 on every actuator (motors included) drives some environments unstable
    within 100 steps.  The environments the engine flags (non-finite) or leaves beyond 1e3 rad/s are
    exactly those whose fp64 oracle rollout blows up; the others stay finite and close (they sit next to the stability limit, where rounding
    differences are amplified: 1e-2 here, against 2e-5 in the stable regime above)."""
    n_envs = 40
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, 'salamander_swim', n_envs, 100, path='fast',
                                                    check_pending=False)
    flagged = set(np.nonzero(physics.flags & 1)[0].tolist())
    qpos, qvel = physics.qpos, physics.qvel
    diverged = set()
    for env in range(n_envs):
        _, _, states = oracle_rollout(spec, model, physics.tables, 101, qpos0[env], qvel0[env], ctrl[env])
        peak = max(float(np.abs(v).max()) for _, v in states)
        if not np.isfinite(peak) or peak > 1e4:
            diverged.add(env)
        else:
            assert np.isfinite(qvel[env]).all() and scaled_error(qpos[env], states[-1][0]) < 1e-2, env
    print('diverged in the oracle:', sorted(diverged), 'flagged by the engine:', sorted(flagged))
    # an environment still on its way up at step 100 is not flagged yet (the flag is 'non-finite')
    blowing_up = {env for env in range(n_envs) if not np.isfinite(qvel[env]).all() or np.abs(qvel[env]).max() > 1e3}
    assert diverged and flagged <= diverged and diverged == flagged | blowing_up


def test_fp32_peak_microbenchmark(cuda_library):
    """fb_measure_fp32_peak: the FFMA roof bench.py quotes the kernels against is a B200-sized
    number (nominal 74.4 TFLOP/s at 1965 MHz) and repeatable."""
    from farms_mujoco_b200.engine import measure_fp32_peak
    a, b = measure_fp32_peak(0, cuda_library), measure_fp32_peak(0, cuda_library)
    assert 40.0 < a < 80.0 and abs(a - b) < 0.1*a, (a, b)


@pytest.mark.parametrize('per_thread', [True, False])
@pytest.mark.parametrize('name', ['swimmer8', 'salamander_swim'])
def test_hand_over_mid_launch(cuda_library, name, per_thread):
    import fastpath_cases
    fastpath_cases.check_hand_over(cuda_library, name, n_envs=70, per_thread=per_thread)


@pytest.mark.parametrize('name', ['salamander', 'centipede'])
def test_constraint_paths_agree(cuda_library, name):
    import fastpath_cases
    # the tolerance is the team kernel's (5e-3 over a ground-contact rollout, see the header): in the
    # worst centipede the per-thread kernel is 2e-5 from the oracle, the team kernel 3e-3
    fastpath_cases.check_constraint_paths_agree(cuda_library, name, n_envs=70, tol=5e-3)


@pytest.mark.parametrize('name', ['salamander_swim', 'salamander'])
def test_block_sizes_agree(cuda_library, name, monkeypatch):
    """16 or 32 environments per warp (fb_create picks 16 for candidate-rich models on small
    batches) run the same per-thread arithmetic: bit-identical results."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, 75)
    outs = []
    for blk in (16, 32):
        monkeypatch.setenv('FARMS_B200_FAST_BLOCK', str(blk))
        physics = BatchedPhysics.from_spec(spec, 75, buffer_size=7, library=cuda_library)
        assert physics.fast_path == blk
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(6)
        outs.append((physics.qpos, physics.qvel, physics.log_arrays()))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and np.array_equal(outs[0][1], other[1])
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.array_equal(outs[0][2][kind], other[2][kind]), kind


@pytest.mark.parametrize('name', ['swimmer8', 'salamander_swim', 'salamander'])
def test_slim_layout_is_bit_identical(cuda_library, name, monkeypatch):
    """Large-batch (SLIM) layout of the unconstrained kernel (1, 4, 7, 8 warps per block; 7 is what
    fb_create picks for 65,536 environments) vs the regular one; the ground model exercises the hand-over out of multi-warp blocks."""
    from farms_mujoco_b200.engine import BatchedPhysics
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', '32')
    spec, model, qpos0, qvel0, ctrl = make_case(name, 75)
    outs = []
    for slim in (0, 1, 4, 7, 8):
        physics = BatchedPhysics.from_spec(spec, 75, buffer_size=9, library=cuda_library)
        if physics.fast_path != 32:
            pytest.skip('SLIM needs 32 environments per warp')
        physics.set_fast_slim(slim)
        assert physics.fast_slim == slim
        # the general variant: one source for every layout, hence the same bits (the LEAN variants
        # are separate compilations that agree to a few ulp, test_lean_variant)
        physics.set_fast_lean(False)
        # ... and one constrained kernel for the hand-overs (its SPLIT variant only exists beside the
        # regular layout and agrees with the single-warp kernel to rounding, test_con_split_variant)
        physics.set_con_split(False)
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


@pytest.mark.parametrize('path,tol', [('fast', 2e-5), ('team', 5e-4)])
def test_box_plane_contacts(cuda_library, path, tol):
    """Plane-box contacts (mjc_PlaneBox corners): SALAMANDER with box feet and a box trunk segment."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_box_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n, n_steps = 70, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=cuda_library)
    physics.set_fast_path(path != 'team')
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.log_arrays()['contacts'].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, [0, 1, 35, n - 1], n_steps, tol,
                                       tol_contacts=max(tol, 5e-4))


def test_constrained_launch_split_is_invariant(cuda_library):
    """Ground contact on the per-thread constrained kernel: 12 steps == 3 launches of 4."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', 40)
    outs = []
    for chunks in ([12], [4, 4, 4]):
        physics = BatchedPhysics.from_spec(spec, 40, buffer_size=13, library=cuda_library)
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        for n in chunks:
            physics.step(n)
        outs.append((physics.qpos, physics.qvel, physics.log_arrays()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert np.array_equal(outs[0][2][kind], outs[1][2][kind]), kind


@pytest.mark.parametrize('variant,free_base', [('swimmer8_features', True), ('swimmer8_fixed_base', False)])
def test_model_variants(cuda_library, variant, free_base):
    import fastpath_cases
    import variant_models
    fastpath_cases.check_variant(cuda_library, getattr(variant_models, variant)(), n_envs=70,
                                 free_base=free_base)


@pytest.mark.parametrize('which', ['swimmer8', 'features', 'salamander'])
def test_ctrl_sequence(cuda_library, which):
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import models
    spec = {'swimmer8': models.swimmer8, 'features': variant_models.swimmer8_features,
            'salamander': models.salamander}[which]()
    fastpath_cases.check_ctrl_sequence(cuda_library, spec, n_envs=70)


def test_paths_agree(cuda_library):
    import fastpath_cases
    fastpath_cases.check_paths_agree(cuda_library, 'salamander_swim', n_envs=96)


@pytest.mark.parametrize('team,name', [(8, 'swimmer8'), (16, 'swimmer8'), (16, 'salamander'),
                                       (32, 'salamander'), (32, 'centipede')])
def test_team_sizes_agree(cuda_library, team, name):
    """Every team width runs the same arithmetic up to reduction order (a team of T lanes
    holds at most 2*T bodies)."""
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, name, 24, 5, team=team)
    assert physics.team_lanes == team
    # contact forces of the team kernel on CENTIPEDE: up to 6e-3 (fp32 CRB + L'DL on 4 g legs)
    _compare(spec, model, physics, qpos0, qvel0, ctrl, [0, 23], 5, 2e-2 if name != 'swimmer8' else 2e-3,
             tol_contacts=5e-2)


def test_launch_split_is_invariant(cuda_library):
    """20 steps in one launch == 4 launches of 5 (state round-trips through HBM)."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander_swim', 16)
    outs = []
    for chunks in ([20], [5, 5, 5, 5]):
        physics = BatchedPhysics.from_spec(spec, 16, buffer_size=21, library=cuda_library)
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        for n in chunks:
            physics.step(n)
        outs.append((physics.qpos, physics.qvel, physics.log_arrays()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert np.array_equal(outs[0][2][kind], outs[1][2][kind]), kind


def test_derived_quantities(cuda_library):
    """FbDerivedView == the oracle's mjData fields for the pre-step state."""
    from oracle.oracle import OraclePhysics
    spec, model, qpos0, qvel0, ctrl, physics = _run(cuda_library, 'salamander', 8, 1)
    der = physics.derived()
    for env in (0, 7):
        orc = OraclePhysics(model)
        orc.reset(keyframe_id=0)
        orc.data.qpos[:] = qpos0[env]
        orc.data.qvel[:] = qvel0[env]
        orc.data.ctrl[:] = ctrl[env]
        orc.forward()
        assert scaled_error(der['xpos'][env], orc.data.xpos) < 1e-6
        assert scaled_error(der['xquat'][env], orc.data.xquat) < 1e-6
        assert scaled_error(der['xipos'][env], orc.data.xipos) < 1e-6
        assert scaled_error(der['linvel'][env], orc.arrays['body_linvel'].reshape(-1, 3)) < 1e-5
        assert scaled_error(der['angvel'][env], orc.arrays['body_angvel'].reshape(-1, 3)) < 1e-5
        assert scaled_error(der['actuator_force'][env], orc.data.actuator_force) < 1e-5
        assert scaled_error(der['qacc'][env], orc.data.qacc) < 1e-4
        assert der['ncon'][env] == orc.ncon
        n = orc.ncon
        assert np.array_equal(der['con_cand'][env][:n], orc.arrays['con_cand'][:n])
        assert scaled_error(der['con_pos'][env][:n], orc.arrays['con_pos'].reshape(-1, 3)[:n]) < 1e-6
        ref_force = orc.arrays['con_force'].reshape(-1, 6)[:n, :3]
        assert scaled_error(der['con_force'][env][:n], ref_force) < 1e-4


def test_full_size_properties(cuda_library):
    """BASELINE size (16,384 swimming salamanders): size-independent properties."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200 import models, mjcf_subset
    spec = models.MODELS['salamander_swim']()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n = 16384
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, size=(1, model.nq - 7))   # same state in every env
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=4, library=cuda_library)
    physics.reset(qpos0, None)
    physics.step(3)
    qpos, qvel = physics.qpos, physics.qvel
    # identical inputs -> bit-identical outputs in every environment
    assert np.array_equal(qpos, np.broadcast_to(qpos[0], qpos.shape))
    assert np.array_equal(qvel, np.broadcast_to(qvel[0], qvel.shape))
    # unit quaternions, finite state, no divergence flag
    assert np.allclose(np.linalg.norm(qpos[:, 3:7], axis=1), 1.0, atol=1e-6)
    assert np.isfinite(qpos).all() and np.isfinite(qvel).all()
    assert not physics.flags.any()
    links = physics.log_arrays(env=n - 1)['links']
    assert np.array_equal(links, physics.log_arrays(env=0)['links'])
    # log quaternions are xyzw unit quaternions; CoM and URDF orientation columns equal
    assert np.allclose(np.linalg.norm(links[:, :, 3:7], axis=-1), 1.0, atol=1e-5)
    assert np.array_equal(links[:, :, 3:7], links[:, :, 10:14])


def test_full_size_ground_properties(cuda_library):
    """BASELINE sizes of the contact configurations (4,096 salamanders / 8,192 centipedes on the
    ground): identical inputs give bit-identical outputs in every environment, the state stays
    finite, and at rest the contact sensors carry the animat's weight."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200 import models, mjcf_subset
    for name, n in (('salamander', 4096), ('centipede', 8192)):
        spec = models.MODELS[name]()
        model = mjcf_subset.parse_mjcf(spec.mjcf)
        rng = np.random.default_rng(11)
        qpos0 = np.tile(model.key_qpos, (n, 1))
        qpos0[:, 7:] += rng.uniform(-0.05, 0.05, size=(1, model.nq - 7))
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=4, library=cuda_library)
        physics.reset(qpos0, None)
        physics.set_ctrl(np.zeros((n, model.nu)))     # position actuators hold the zero posture
        for _ in range(25):
            physics.step(16)
        assert physics.constraint_path == 1 and physics.last_pending == n
        qpos, qvel = physics.qpos, physics.qvel
        assert not physics.flags.any()
        assert np.isfinite(qpos).all() and np.isfinite(qvel).all()
        assert np.array_equal(qpos, np.broadcast_to(qpos[0], qpos.shape))
        assert np.array_equal(qvel, np.broadcast_to(qvel[0], qvel.shape))
        contacts = physics.log_arrays(env=n - 1)['contacts']
        assert np.array_equal(contacts, physics.log_arrays(env=0)['contacts'])
        # settled after 0.4 s: the summed vertical reaction + friction balances the weight
        total = contacts[-1, :, 6:9].sum(axis=0)
        weight = float(np.sum(model.body_mass))*9.81
        assert abs(abs(total[2]) - weight) < 0.1*weight, (total, weight)


def test_lanes_are_independent(cuda_library):
    """A warp mixes environments on the ground with environments held in the air: every
    environment gets bit for bit what it gets in a batch of its own kind (the constrained step
    votes across the warp only to SKIP work, never to change a lane's arithmetic)."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', 64, qvel_scale=0.05)
    high = qpos0.copy()
    high[:, 2] += 0.5                       # free fall, no contact within the rollout
    mixed = qpos0.copy()
    mixed[1::2] = high[1::2]
    outs = {}
    for key, q in (('ground', qpos0), ('air', high), ('mixed', mixed)):
        physics = BatchedPhysics.from_spec(spec, 64, buffer_size=9, library=cuda_library)
        # the single-warp constrained kernel in all three batches: the all-ground one would otherwise
        # run its SPLIT variant, which agrees to rounding only (test_con_split_variant)
        physics.set_con_split(False)
        physics.reset(q, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(8)
        outs[key] = (physics.qpos, physics.qvel, physics.log_arrays())
    for env in range(64):
        ref = outs['air'] if env % 2 else outs['ground']
        assert np.array_equal(outs['mixed'][0][env], ref[0][env]), env
        assert np.array_equal(outs['mixed'][1][env], ref[1][env]), env
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.array_equal(outs['mixed'][2][kind][env], ref[2][kind][env]), (env, kind)
    assert outs['ground'][2]['contacts'].any() and not outs['air'][2]['contacts'].any()


def test_simulation_layer(cuda_library):
    """Reference-facing loop (Simulation.run, fused launches, on-device controller) on the GPU
    vs the oracle's replay of the reference's Simulation.run."""
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.control import TravellingWaveController
    from farms_mujoco_b200.models import travelling_wave_parameters
    from farms_mujoco_b200.simulation.simulation import Simulation
    n_it, n_envs = 24, 96
    spec = models.salamander(swimming=True, n_iterations=n_it)
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    phase = np.linspace(0.0, 3.0, n_envs)
    sim = Simulation.from_spec(spec, n_envs=n_envs, chunk=8, library=cuda_library,
                               controller=TravellingWaveController(joints, amp, freq, lag, env_phase=phase))
    sim.run()
    assert sim.task.device_controller and sim.iteration == n_it - 1
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    for env in (0, 50, n_envs - 1):
        def controller(iteration, time, env=env):
            ctrl = np.zeros(model.nu)
            ctrl[acts] = amp*np.sin(2*np.pi*freq*time - lag + phase[env])
            return ctrl
        data, _ = fo.reference_rollout(OraclePhysics(model), spec, sim.physics.tables, n_it,
                                       controller=controller)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            ours = getattr(sim.task.data.sensors, kind).array[env]
            assert log_error(kind, ours, getattr(data.sensors, kind).array) < 5e-5, kind


def test_step_host_joint_columns(cuda_library):
    """fb_set_host_joint_columns: the joints row comes down as the selected columns."""
    import torch
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.layout import sc
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', 70)
    nl, nj = len(spec.links_names), len(spec.joints_names)
    cols = [sc.joint_position, sc.joint_velocity, sc.joint_torque, sc.joint_limit_force]
    rows = {}
    for compact in (False, True):
        physics = BatchedPhysics.from_spec(spec, 70, buffer_size=8, library=cuda_library)
        physics.reset(qpos0, qvel0)
        links = torch.zeros((70, nl, 20), dtype=torch.float32).pin_memory()
        joints = torch.zeros((70, nj, len(cols) if compact else sc.joint_size), dtype=torch.float32).pin_memory()
        if compact:
            physics.set_host_joint_columns(cols)
        physics.step_host(4, ctrl=torch.as_tensor(ctrl, dtype=torch.float32).pin_memory(), links_row=links,
                          joints_row=joints)
        rows[compact] = (links.clone(), joints.clone())
    assert torch.equal(rows[True][0], rows[False][0])
    assert torch.equal(rows[True][1], rows[False][1][:, :, cols])
    assert rows[True][1][:, :, :3].abs().sum() > 0


def test_step_host_pipelined_equals_synchronous(cuda_library):
    """fb_step_host_async (download of call i overlapping the kernels of call i+1, two host
    buffer sets) returns the rows fb_step_host returns."""
    import torch
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander_swim', 64)
    nl, nj = len(spec.links_names), len(spec.joints_names)
    rows = {}
    base = torch.as_tensor(ctrl, dtype=torch.float32)
    for mode, nsets in (('sync', 2), ('pipelined', 2), ('pipelined3', 3)):
        physics = BatchedPhysics.from_spec(spec, 64, buffer_size=8, library=cuda_library)
        physics.reset(qpos0, qvel0)
        ctrl_host = [base.clone().pin_memory() for _ in range(nsets)]
        links = [torch.zeros((64, nl, 20), dtype=torch.float32).pin_memory() for _ in range(nsets)]
        joints = [torch.zeros((64, nj, 18), dtype=torch.float32).pin_memory() for _ in range(nsets)]
        got, pending = [], [None]*nsets
        for call in range(7):
            k = call % nsets
            if mode != 'sync' and pending[k] is not None:
                if nsets == 2:
                    physics.host_wait_slot(k)
                else:
                    physics.host_wait_call(pending[k])
                got.append((links[k].clone(), joints[k].clone()))
            ctrl_host[k].copy_(base*(1.0 - 0.1*call))          # a new ctrl per call (pinned: SM upload path)
            pending[k] = physics.step_host(3, ctrl=ctrl_host[k], links_row=links[k], joints_row=joints[k],
                                           pipelined=mode != 'sync')
            assert pending[k] == call
            if mode == 'sync':
                got.append((links[k].clone(), joints[k].clone()))
        if mode != 'sync':
            physics.host_wait()
            for call in range(7 - nsets, 7):
                got.append((links[call % nsets].clone(), joints[call % nsets].clone()))
        rows[mode] = got
    assert len(rows['sync']) == len(rows['pipelined']) == len(rows['pipelined3']) == 7
    for other in ('pipelined', 'pipelined3'):
        for (la, ja), (lb, jb) in zip(rows['sync'], rows[other]):
            assert torch.equal(la, lb) and torch.equal(ja, jb)
    assert rows['sync'][-1][0].abs().sum() > 0
    assert not torch.equal(rows['sync'][-1][1], rows['sync'][-2][1])
    # pageable ctrl (plain numpy): the cudaMemcpyAsync path gives the same rows
    physics = BatchedPhysics.from_spec(spec, 64, buffer_size=8, library=cuda_library)
    physics.reset(qpos0, qvel0)
    lrow, jrow = np.zeros((64, nl, 20), np.float32), np.zeros((64, nj, 18), np.float32)
    for call in range(7):
        physics.step_host(3, ctrl=np.ascontiguousarray((base*(1.0 - 0.1*call)).numpy()), links_row=lrow, joints_row=jrow)
    assert np.array_equal(lrow, rows['sync'][-1][0].numpy()) and np.array_equal(jrow, rows['sync'][-1][1].numpy())


@pytest.mark.parametrize('which', ['features', 'fixed_base'])
def test_slim_layout_variants_and_ctrl_sequence(cuda_library, which, monkeypatch):
    """SLIM layout on the hand-edited models (slide joint, off-origin anchors, clamps, fixed base)
    with an uploaded control sequence: bit-identical to the regular layout."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', '32')
    spec = variant_models.swimmer8_features() if which == 'features' else variant_models.swimmer8_fixed_base()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n, n_steps = 75, 7
    rng = np.random.default_rng(2)
    first = 7 if which == 'features' else 0
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, first:] += rng.uniform(-0.1, 0.1, (n, model.nq - first))
    qvel0 = rng.uniform(-0.3, 0.3, (n, model.nv))
    seq = rng.uniform(-0.3, 0.3, (n_steps, n, model.nu)).astype(np.float32)
    outs = []
    for slim in (0, 1, 7):
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=cuda_library)
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


# ---- the BASELINE.json configurations themselves (batch size, layout and block size fb_create
# picks for them, a ring that wraps), sampled environments against the oracle -------------------
BENCH_CONFIGS = [
    # name, envs, ring, launches of 16 steps, tol (state, links / joints / xfrc rows), tol contacts
    ('salamander_swim', 65536, 64, 6, 5e-5, 5e-5),      # configs[4] at N = 1: SLIM layout, 7 warps per block, ring wraps
    ('salamander_swim', 16384, 64, 6, 5e-5, 5e-5),      # configs[2]: regular layout
    ('salamander_swim', 8192, 64, 6, 5e-5, 5e-5),       # configs[4] at N = 8 (65,536 / 8 per GPU)
    ('swimmer8', 65536, 64, 6, 5e-5, 5e-5),
    ('salamander', 4096, 16, 2, 1e-4, 5e-4),            # configs[1]: ground contact in every step
    ('centipede', 8192, 16, 2, 1e-4, 5e-4),             # configs[3]
]


@pytest.mark.parametrize('name,n_envs,ring,launches,tol,tol_contacts', BENCH_CONFIGS)
def test_bench_configuration_against_oracle(cuda_library, name, n_envs, ring, launches, tol, tol_contacts):
    """bench.py's own setup (synthetic_inputs, on-device travelling wave, 16 steps per launch, the
    kernel variant fb_create selects for the batch) at BASELINE.json's sizes: 16 sampled
    environments -- first / last of the first / last block and warp, plus random ones -- agree with
    the fp64 oracle over a rollout that wraps the ring."""
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.models import travelling_wave_parameters
    from farms_mujoco_b200.sharding import synthetic_inputs
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    from conftest import log_errors, state_errors
    inner = 16
    spec = models.MODELS[name]()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    qpos0, qvel0, phase = synthetic_inputs(model, np.arange(n_envs))
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=ring, library=cuda_library)
    physics.set_env_phase(phase)
    physics.set_wave_controller(acts, amp, freq, lag)
    physics.reset(qpos0, qvel0)
    if name == 'salamander_swim' and n_envs == 65536:
        assert physics.fast_slim == 7, physics.fast_slim      # the variant the bench line is measured on
    if name == 'salamander_swim' and n_envs <= 16384:
        assert physics.fast_slim == 0
    for _ in range(launches):
        physics.step(inner, sync=False)
    physics.synchronize()
    n_steps = inner*launches
    assert n_steps > ring
    swimming = name in SWIMMING
    assert physics.last_pending == (0 if swimming else n_envs)
    assert not physics.flags.any()
    rng = np.random.default_rng(5)
    wpb = max(1, physics.fast_slim)*32
    sample = {0, 31, 32, wpb - 1, wpb, n_envs - wpb, n_envs - 33, n_envs - 32, n_envs - 1, n_envs//2}
    sample |= set(int(e) for e in rng.integers(0, n_envs, size=16 - len(sample)))
    qpos, qvel = physics.qpos, physics.qvel
    worst = {}
    for env in sorted(sample):
        def controller(iteration, time, env=env):
            ctrl = np.zeros(model.nu)
            ctrl[acts] = amp*np.sin(2*np.pi*freq*time - lag + phase[env])
            return ctrl
        data, states = fo.reference_rollout(OraclePhysics(model), spec, physics.tables, n_steps + 1,
                                            controller=controller, qpos0=qpos0[env], qvel0=qvel0[env])
        ours = physics.export_farms(env)
        for key, val in state_errors(qpos[env], qvel[env], *states[-1]).items():
            worst[key] = max(worst.get(key, 0.0), val)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            ref = getattr(data.sensors, kind).array
            want = np.zeros((ring,) + ref.shape[1:])
            for it in range(n_steps + 1):
                want[it % ring] = ref[it]
            for key, val in log_errors(kind, getattr(ours.sensors, kind).array, want).items():
                worst[f'{kind}.{key}'] = max(worst.get(f'{kind}.{key}', 0.0), val)
    print(name, n_envs, 'envs,', n_steps, 'steps:', {k: f'{v:.1e}' for k, v in worst.items() if v > 0})
    for key, val in worst.items():
        assert val < (tol_contacts if key.startswith('contacts') else tol), (key, val, worst)


@pytest.mark.parametrize('kind', ['limits', 'contacts'])
def test_ring_wrap(cuda_library, kind):
    """buffer_size < n_steps, constraint columns dirtied early and clean later."""
    import fastpath_cases
    fastpath_cases.check_ring_wrap(cuda_library, kind, n_envs=70)


def test_reset_clears_constraint_columns(cuda_library):
    import fastpath_cases
    fastpath_cases.check_reset_clears_log(cuda_library, n_envs=70)


def test_device_cpg_matches_host_controller(cuda_library):
    """the on-device CPG (fb_cpg_kernel feeding the control sequence of every launch)
    against the same network stepped on the host through ExperimentTask.step_control."""
    import fastpath_cases
    fastpath_cases.check_device_cpg(cuda_library, n_envs=70, n_it=48, chunk=16)


def test_lean_variant(cuda_library, monkeypatch):
    import fastpath_cases
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', '32')
    fastpath_cases.check_lean_variant(cuda_library, n_envs=75, slims=(0, 1, 7))


def test_split_variant_is_bit_identical(cuda_library, monkeypatch):
    """Small-batch SPLIT variant of the unconstrained kernel (fb_fast_split_kernel)."""
    import fastpath_cases
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', '32')
    fastpath_cases.check_split_variant_is_bit_identical(cuda_library)


@pytest.mark.parametrize('block', ['16', '32'])
def test_con_split_variant(cuda_library, monkeypatch, block):
    """SPLIT variant of the constrained kernel (fb_fastc_split_kernel), 16 and 32 environments per group."""
    import fastpath_cases
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', block)
    fastpath_cases.check_con_split_variant(cuda_library)


def test_split_variant_fixed_base(cuda_library, monkeypatch):
    """Tree-split kernel on a branching tree WITHOUT a floating root (the warps other than the trunk's
    have no published root position to read): against the oracle and against the single-warp kernel."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', '32')
    spec = variant_models.salamander_swim_fixed_base()
    fastpath_cases.check_variant(cuda_library, spec, free_base=False)
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n = 70
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1)) + rng.uniform(-0.1, 0.1, (n, model.nq))
    qvel0 = rng.uniform(-0.3, 0.3, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    outs = []
    for split in (False, True):
        physics = BatchedPhysics.from_spec(spec, n, buffer_size=9, library=cuda_library)
        physics.set_fast_split(split)
        assert bool(physics.fast_split) == split
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(8)
        outs.append((physics.qpos, physics.qvel, physics.log_arrays()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert np.array_equal(outs[0][2][kind], outs[1][2][kind]), kind


@pytest.mark.parametrize('block', ['16', '32'])
def test_con_split_mixed_groups(cuda_library, monkeypatch, block):
    """Groups handed over in full go to the SPLIT constrained kernel, late arrivals of the same launch
    to the single-warp kernel behind it (the per-group flags of fb_fastc_split_kernel)."""
    import fastpath_cases
    monkeypatch.setenv('FARMS_B200_FAST_BLOCK', block)
    fastpath_cases.check_con_split_mixed_groups(cuda_library)


@pytest.mark.parametrize('path,tol', [('fast', 2e-5), ('fast_single_warp', 2e-5), ('team', 1e-3)])
def test_cylinder_plane_contacts(cuda_library, path, tol):
    """Plane-cylinder contacts (mjc_PlaneCylinder: up to four points per cylinder): SALAMANDER with
    tilted cylinder feet and a cylinder trunk segment, against the oracle."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_cylinder_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n, n_steps = 70, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=cuda_library)
    physics.set_fast_path(path != 'team')
    physics.set_con_split(path == 'fast')
    assert not physics.fast_lean
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.log_arrays()['contacts'][:, :, 12:].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, [0, 1, 35, n - 1], n_steps, tol,
                                       tol_contacts=max(tol, 5e-4))


@pytest.mark.parametrize('path,tol', [('fast', 2e-5), ('fast_single_warp', 2e-5), ('team', 5e-4)])
def test_ellipsoid_plane_contacts(cuda_library, path, tol):
    """Plane-ellipsoid contacts (mjc_PlaneConvex: the support point along the plane normal): SALAMANDER
    with rotated ellipsoid feet and an ellipsoid trunk segment, against the oracle -- SPLIT and
    single-warp constrained kernels (general variants: the LEAN ones leave the branch out), team kernel."""
    import fastpath_cases
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    spec = variant_models.salamander_ellipsoid_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n, n_steps = 70, 10
    rng = np.random.default_rng(3)
    qpos0 = np.tile(model.key_qpos, (n, 1))
    qpos0[:, 7:] += rng.uniform(-0.1, 0.1, (n, model.nq - 7))
    qvel0 = rng.uniform(-0.2, 0.2, (n, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n, model.nu))
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=cuda_library)
    physics.set_fast_path(path != 'team')
    physics.set_con_split(path == 'fast')
    assert not physics.fast_lean
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.log_arrays()['contacts'][:, :, 12:].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, [0, 1, 35, n - 1], n_steps, tol,
                                       tol_contacts=max(tol, 5e-4))


@pytest.mark.gpu
def test_drag_operator(cuda_library):
    """fb_drag_forces on the device (float64) against the oracle's drag_forces: 1e-12 relative."""
    from drag_cases import check_operator
    assert check_operator(cuda_library, n=4099) < 1e-12


@pytest.mark.gpu
def test_sub_steps_log_full_steps_only(cuda_library):
    """num_sub_steps = 2 through the Simulation layer on the device: host rows = the reference's
    full-step rows (device row 2i), against the oracle's literal replay of the sub-stepped loop."""
    from test_simulation_loop import _substep_sim, _substep_check, HostWave
    from farms_mujoco_b200 import models
    from farms_mujoco_b200.models import travelling_wave_parameters
    n_it, phase = 8, [0.3, 1.1, 2.0]
    spec = models.salamander(swimming=True, n_iterations=n_it)
    wave = travelling_wave_parameters(spec)
    sim = _substep_sim(spec, 2, cuda_library, n_envs=3, controller=HostWave(*wave, phase))
    sim.run()
    assert sim.physics.iteration == 2*n_it - 1 and sim.iteration == n_it
    _substep_check(sim, spec, n_it, 2, phase, wave)


@pytest.mark.gpu
def test_pair_self_collisions(cuda_library):
    """Explicit <contact><pair> sphere-sphere self-collisions on the device: the team kernel with the
    dense Newton Hessian (rows on two branches of the tree), against the oracle."""
    import fastpath_cases
    from test_emu_parity import _pair_case
    from farms_mujoco_b200.engine import BatchedPhysics
    n, n_steps = 40, 12
    spec, model, qpos0, qvel0, ctrl = _pair_case(n)
    physics = BatchedPhysics.from_spec(spec, n, buffer_size=n_steps + 1, library=cuda_library)
    assert physics.fast_path == 0 and physics.team_lanes > 1
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    contacts = physics.log_arrays()['contacts']
    names = [tuple(c) for c in spec.contacts_names]
    pair = names.index(('link_leg_0_L_3', 'link_leg_0_R_3'))
    assert np.abs(contacts[0, 1:, pair, 6:9]).max() > 1e-3 and not contacts[n - 1, :, pair].any()
    fastpath_cases.compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, (0, 1, 17, n - 2, n - 1), n_steps, 2e-4,
                                       tol_contacts=2e-4)


@pytest.mark.gpu
def test_torque_control_disables_position_actuators(cuda_library):
    """initialize_control's model edit (task.py:262-286), torque commands and spring references
    through the Simulation layer on the device, against the oracle on the edited model."""
    from test_simulation_loop import _torque_control_case
    _torque_control_case(cuda_library)
