"""Cases shared by the emulation (CPU) and GPU parity tests of the two-kernel step:
environment-per-thread kernel first, team kernel on what it hands over."""

import numpy as np

from conftest import log_errors, make_case, oracle_rollout, scaled_error, parity_errors, log_error, state_errors


def compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, envs, n_steps, tol, tol_contacts=None):
    qpos, qvel, logs = physics.qpos, physics.qvel, physics.log_arrays()
    # bit 2 (solver stopped on its iteration cap in fp32) is informational: the comparison
    # with the oracle below is the check; non-finite state / contact overflow are errors
    # (checked on the compared environments: a random scenario can be physically unstable -- the
    # fp64 oracle blows up too -- and which step first exceeds 1e30 then depends on rounding)
    assert not (physics.flags[list(envs)] & 3).any(), physics.flags
    assert np.count_nonzero(physics.flags & 3) <= len(physics.flags)//50, physics.flags
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
    return worst


def hand_over_case(name, n_envs):
    """Joints start just inside their upper limit; the position actuator of a driven joint pulls it
    through the limit after a few steps (a different number in every environment, some never)."""
    spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs, qvel_scale=0.0, ctrl_scale=0.0)
    rng = np.random.default_rng(5)
    hi = np.asarray(model.jnt_range).reshape(-1, 2)[:, 1]
    gain = np.asarray(model.actuator_gainprm).reshape(model.nu, -1)[:, 0]
    bias = np.asarray(model.actuator_biasprm).reshape(model.nu, -1)
    trn = np.asarray(model.actuator_trnid).reshape(model.nu, -1)[:, 0]
    position_act = {int(trn[a]): a for a in range(model.nu) if gain[a] != 0 and bias[a, 1] == -gain[a]}
    qpos0[:, 7:] = 0.0
    qvel0[:] = 0.0
    ctrl[:] = 0.0
    # joint 0 is the free joint; the salamander's leg joints (12 ..) carry 2 g links and go
    # unstable when driven through a limit at this time step (the fp64 oracle blows up too)
    joint = rng.integers(1, min(model.njnt, 12), size=n_envs)
    driven = rng.uniform(size=n_envs) < 0.7
    driven[0], driven[-1] = True, False
    rows = np.arange(n_envs)
    qadr = np.asarray(model.jnt_qposadr)[joint]
    # start a little inside the upper limit; the position actuator of a driven joint pulls
    # it through the limit after a few steps (a different number in every environment)
    qpos0[rows, qadr] = hi[joint] - rng.uniform(0.0, 0.004, size=n_envs)
    for e in range(n_envs):
        ctrl[e, position_act[int(joint[e])]] = hi[joint[e]] + (0.6 if driven[e] else -0.2)
    return spec, model, qpos0, qvel0, ctrl


def check_hand_over(library, name, n_envs, n_steps=16, tol=2e-3, per_thread=True):
    """Joints start just inside their upper limit and move into it: the per-thread kernel
    takes the first steps, hands an environment over when its limit becomes active (a
    different step in every environment, some never), and the team kernel finishes the
    launch with the limit row in the solver.  The log must be the oracle's throughout."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.layout import sc
    spec, model, qpos0, qvel0, ctrl = hand_over_case(name, n_envs)
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=library)
    assert physics.fast_path
    physics.set_constraint_path(per_thread)
    if per_thread:
        tol = min(tol, 5e-5)
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    pending = physics.last_pending
    assert 0 < pending < n_envs, pending
    envs = sorted({0, 1, n_envs//2, n_envs - 2, n_envs - 1})
    compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, envs, n_steps, tol)
    limit_force = physics.log_arrays()['joints'][:, :, :, sc.joint_limit_force]
    assert (np.abs(limit_force).max(axis=(1, 2)) > 0).sum() == pending


def check_paths_agree(library, name, n_envs, n_steps=10, tol=1e-4):
    """Per-thread kernel (ABA) vs team kernel (CRB + L'DL) on the same unconstrained rollout."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs)
    outs = {}
    for fast in (True, False):
        physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=library)
        physics.set_fast_path(fast)
        assert bool(physics.fast_path) == fast
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(n_steps)
        assert physics.last_pending == (0 if fast else n_envs)
        outs[fast] = (physics.qpos, physics.qvel, physics.log_arrays(), physics.xfrc_applied)
    assert scaled_error(outs[True][0], outs[False][0]) < tol
    assert scaled_error(outs[True][1], outs[False][1]) < tol
    assert scaled_error(outs[True][3], outs[False][3]) < tol
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert scaled_error(outs[True][2][kind], outs[False][2][kind]) < tol, kind


def check_constraint_paths_agree(library, name, n_envs, n_steps=10, tol=2e-4):
    """Per-thread constrained kernel (matrix-free Newton on the ABA) vs team kernel (CRB + L'DL +
    Newton in M's sparse layout) on the same rollout with ground contact."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs)
    outs = {}
    for per_thread in (True, False):
        physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=library)
        physics.set_constraint_path(per_thread)
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        physics.step(n_steps)
        assert physics.last_pending == n_envs
        assert not (physics.flags & 3).any()
        outs[per_thread] = (physics.qpos, physics.qvel, physics.log_arrays())
    assert outs[True][2]['contacts'].any()
    assert scaled_error(outs[True][0], outs[False][0]) < tol
    assert scaled_error(outs[True][1], outs[False][1]) < tol
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert scaled_error(outs[True][2][kind], outs[False][2][kind]) < tol, kind


def check_variant(library, spec, n_envs=3, n_steps=15, tol=2e-5, free_base=True):
    """Per-thread kernel on a hand-edited model (tests/variant_models.py) vs the oracle."""
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    rng = np.random.default_rng(1)
    qpos0 = np.tile(model.key_qpos, (n_envs, 1))
    first = 7 if free_base else 0
    qpos0[:, first:] += rng.uniform(-0.1, 0.1, (n_envs, model.nq - first))
    qvel0 = rng.uniform(-0.3, 0.3, (n_envs, model.nv))
    ctrl = rng.uniform(-0.3, 0.3, (n_envs, model.nu))
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=library)
    assert physics.fast_path
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(n_steps)
    assert physics.last_pending == 0
    envs = sorted({0, n_envs//2, n_envs - 1})
    return compare_with_oracle(spec, model, physics, qpos0, qvel0, ctrl, envs, n_steps, tol)


def check_ctrl_sequence(library, spec, n_envs=4, n_steps=9, free_base=True):
    """fb_set_ctrl_sequence: K steps in fused launches == K launches with set_ctrl before each."""
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics, EngineError
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    rng = np.random.default_rng(7)
    qpos0 = np.tile(model.key_qpos, (n_envs, 1))
    first = 7 if free_base else 0
    qpos0[:, first:] += rng.uniform(-0.1, 0.1, (n_envs, model.nq - first))
    qvel0 = rng.uniform(-0.2, 0.2, (n_envs, model.nv))
    seq = rng.uniform(-0.3, 0.3, (n_steps, n_envs, model.nu)).astype(np.float32)
    outs = []
    for mode in ('stepwise', 'fused'):
        physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=n_steps + 1, library=library)
        # launches that read a control sequence use the general kernel variant; bit-identity with
        # the step-by-step run holds within that variant (the LEAN one agrees to a few ulp)
        physics.set_fast_lean(False)
        physics.reset(qpos0, qvel0)
        if mode == 'stepwise':
            for k in range(n_steps):
                physics.set_ctrl(seq[k])
                physics.step(1)
        else:
            physics.set_ctrl_sequence(seq)
            physics.step(4)
            physics.step(n_steps - 4)
            try:
                physics.set_ctrl_sequence(seq[:2])
                physics.step(3)
                raise AssertionError('a launch longer than the sequence must fail')
            except EngineError as err:
                assert 'sequence' in str(err)
            physics.set_ctrl_sequence(None)
        outs.append((physics.qpos, physics.qvel, physics.ctrl, physics.log_arrays()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])          # ctrl ends at the last entry
    for kind in ('links', 'joints', 'contacts', 'xfrc'):
        assert np.array_equal(outs[0][3][kind], outs[1][3][kind]), kind


def _wrapped(reference, ring, last):
    """Rows of a [n_rows, ...] log as a ring of `ring` rows holds them after iteration `last`:
    row r = the latest iteration <= last with iteration % ring == r (task.py:156-166)."""
    out = np.zeros((ring,) + reference.shape[1:])
    for it in range(last + 1):
        out[it % ring] = reference[it]
    return out


def check_ring_wrap(library, kind, n_envs, ring=8, chunks=(3, 5, 7, 2, 8, 6, 9), tol=2e-4, tol_contacts=5e-4):
    """buffer_size < n_steps (update_sensors' ring index, task.py:156-166): the device ring holds
    the latest `ring` iterations, and constraint-only columns that were non-zero on an earlier
    visit of a ring row read zero again once the environment is unconstrained (the
    unconstrained kernel zero-fills exactly the rows a constrained step may have dirtied).

    kind 'limits': swimmers driven into a joint limit for the first steps, then pulled back.
    kind 'contacts': salamanders that start in ground contact with an upward velocity and leave
    the ground after a few steps."""
    from farms_mujoco_b200.engine import BatchedPhysics
    from farms_mujoco_b200.layout import sc
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    n_steps = int(sum(chunks))
    if kind == 'limits':
        spec, model, qpos0, qvel0, ctrl = make_case('swimmer8', n_envs, qvel_scale=0.0, ctrl_scale=0.0)
        hi = np.asarray(model.jnt_range).reshape(-1, 2)[:, 1]
        gain = np.asarray(model.actuator_gainprm).reshape(model.nu, -1)[:, 0]
        bias = np.asarray(model.actuator_biasprm).reshape(model.nu, -1)
        trn = np.asarray(model.actuator_trnid).reshape(model.nu, -1)[:, 0]
        position_act = {int(trn[a]): a for a in range(model.nu) if gain[a] != 0 and bias[a, 1] == -gain[a]}
        rng = np.random.default_rng(11)
        joint = rng.integers(1, model.njnt, size=n_envs)
        qadr = np.asarray(model.jnt_qposadr)[joint]
        qpos0[:, 7:] = 0.0
        qvel0[:] = 0.0
        qpos0[np.arange(n_envs), qadr] = hi[joint] - rng.uniform(0.0, 0.003, size=n_envs)
        ctrl_a, ctrl_b = np.zeros_like(ctrl), np.zeros_like(ctrl)
        for e in range(n_envs):
            ctrl_a[e, position_act[int(joint[e])]] = hi[joint[e]] + 0.6     # into the limit ...
            ctrl_b[e, position_act[int(joint[e])]] = hi[joint[e]] - 0.5     # ... and back out
        switch = chunks[0] + chunks[1]
    else:
        spec, model, qpos0, qvel0, ctrl = make_case('salamander', n_envs, qvel_scale=0.05, ctrl_scale=0.1)
        qvel0[:, 2] = 1.2                                   # leaves the ground within a few steps
        ctrl_a = ctrl_b = ctrl
        switch = n_steps
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=ring, library=library)
    assert physics.fast_path and physics.constraint_path == 1
    physics.reset(qpos0, qvel0)
    done = 0
    for n in chunks:
        physics.set_ctrl(ctrl_a if done < switch else ctrl_b)
        physics.step(n)
        done += n
    logs = physics.log_arrays()
    dirty_col = sc.joint_limit_force
    saw_nonzero = False
    for env in sorted({0, 1, n_envs//2, n_envs - 1}):
        data, states = fo.reference_rollout(
            OraclePhysics(model), spec, physics.tables, n_steps + 1,
            controller=lambda it, t, e=env: ctrl_a[e] if it < switch else ctrl_b[e],
            qpos0=qpos0[env], qvel0=qvel0[env])
        errs = state_errors(physics.qpos[env], physics.qvel[env], *states[-1])
        assert max(errs.values()) < tol, errs
        for name in ('links', 'joints', 'contacts', 'xfrc'):
            ref = getattr(data.sensors, name).array
            want = _wrapped(ref, ring, n_steps)
            err = log_error(name, logs[name][env], want)
            assert err < (tol_contacts if name == 'contacts' else tol), (name, env, err)
        # the scenario is what the docstring says: non-zero constraint columns early on, all of
        # them zero in the rows the ring holds at the end
        if kind == 'limits':
            early, late = data.sensors.joints.array[:ring, :, dirty_col], logs['joints'][env][:, :, dirty_col]
        else:
            early, late = data.sensors.contacts.array[:ring], logs['contacts'][env]
        saw_nonzero |= bool(np.abs(early).max() > 0)
        assert not np.abs(late).max() > 0, (kind, env)
    assert saw_nonzero


def check_reset_clears_log(library, n_envs=4, ring=8):
    """A second episode on the same handle: rows the first episode dirtied (ground contact) and
    that the second never reaches with a constrained step must read as the oracle's second
    rollout, i.e. zero contacts (fb_reset clears the constraint-only columns of the log)."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', n_envs, qvel_scale=0.05, ctrl_scale=0.1)
    physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=ring, library=library)
    physics.reset(qpos0, qvel0)
    physics.set_ctrl(ctrl)
    physics.step(ring + 3)
    assert physics.log_arrays()['contacts'].any()
    high = qpos0.copy()
    high[:, 2] += 0.5                                       # free fall: no contact in this episode
    physics.reset(high, qvel0)
    physics.set_ctrl(ctrl)
    n_steps = ring - 3
    physics.step(n_steps)
    assert physics.last_pending == 0
    logs = physics.log_arrays()
    assert not logs['contacts'].any()
    for env in (0, n_envs - 1):
        _, data, states = oracle_rollout(spec, model, physics.tables, n_steps + 1, high[env], qvel0[env], ctrl[env])
        assert max(state_errors(physics.qpos[env], physics.qvel[env], *states[-1]).values()) < 2e-5
        for name in ('links', 'joints', 'contacts', 'xfrc'):
            ref = getattr(data.sensors, name).array
            assert log_error(name, logs[name][env][:n_steps + 1], ref) < 2e-5, name
        # rows the second episode has not reached: the first episode's contacts / limit forces are gone
        assert not logs['contacts'][env][n_steps + 1:].any()
        assert not logs['joints'][env][n_steps + 1:].any()


def swimmer_cpg(n_envs, device, torque_joints=(), spring_joints=()):
    """A salamander-type CPG for SWIMMER8: two oscillators per joint (left / right, antiphase),
    nearest-neighbour couplings with a head-to-tail phase lag, position targets = r_L (1 + cos) -
    r_R (1 + cos); `torque_joints` are driven through their motor actuators instead, and the spring
    references of `spring_joints` follow the network too."""
    from farms_mujoco_b200.control import CPGController, ControlType
    nj = 7
    freq, amp, rate = np.full(2*nj, 1.5), np.full(2*nj, 0.15), np.full(2*nj, 20.0)
    couplings = []
    for j in range(nj):
        couplings += [(2*j, 2*j + 1, 10.0, np.pi), (2*j + 1, 2*j, 10.0, np.pi)]
        if j + 1 < nj:
            lag = 2*np.pi/nj
            for s in (0, 1):
                couplings += [(2*j + s, 2*(j + 1) + s, 10.0, lag), (2*(j + 1) + s, 2*j + s, 10.0, -lag)]
    outputs = []
    for j in range(nj):
        name = f'joint_{j}'
        if name in torque_joints:
            outputs.append((name, ControlType.TORQUE, 2*j, -1, 0.02, 0.0))
        else:
            outputs.append((name, ControlType.POSITION, 2*j, 2*j + 1, 1.0, 0.0))
    rng = np.random.default_rng(4)
    phase0 = rng.uniform(0, 2*np.pi, (n_envs, 2*nj))
    springs = [(name, 2*int(name.split('_')[1]), 2*int(name.split('_')[1]) + 1, 2.0, 0.05*(k + 1))
               for k, name in enumerate(spring_joints)]
    return CPGController(freq, amp, rate, couplings, outputs, phase0, amplitude0=0.05, device=device, springs=springs)


def check_device_cpg(library, n_envs=3, n_it=40, chunk=8, tol=2e-5):
    """On-device CPG (fb_set_cpg; position targets, torque commands and spring references) vs the
    same network evaluated on the host through ExperimentTask.step_control (task.py:288-346),
    iteration by iteration."""
    import dataclasses
    import re
    from farms_mujoco_b200 import models
    from farms_mujoco_b200.simulation.simulation import Simulation
    torque_joints = ('joint_5', 'joint_6')
    logs, springs = {}, {}
    for device in (True, False):
        spec = models.swimmer8(n_iterations=n_it)
        # the torque-controlled joints are elastic, and their spring references follow the network
        mjcf, count = re.subn(r'(<joint name="joint_[56]"[^>]*?)stiffness="0.0"', r'\1stiffness="0.05"', spec.mjcf)
        assert count == 2
        spec = dataclasses.replace(spec, mjcf=mjcf)
        # joints 5 and 6 are torque-controlled: their motors lose the 'position' control type, so
        # initialize_control switches their position / velocity actuators off (task.py:274-286)
        for motor in spec.animat_options.control.motors:
            if motor['joint_name'] in torque_joints:
                motor.control_types = ['torque']
        sim = Simulation.from_spec(spec, n_envs=n_envs, chunk=chunk, library=library,
                                   controller=swimmer_cpg(n_envs, device, torque_joints, spring_joints=torque_joints))
        sim.run()
        assert sim.task.device_controller == device and sim.iteration == n_it - 1
        logs[device] = {k: getattr(sim.task.data.sensors, k).array.copy() for k in ('links', 'joints', 'xfrc')}
        springs[device] = sim.physics.qpos_spring
        if device:
            phase, amplitude = sim.physics.cpg_state()
            assert np.isfinite(phase).all() and (amplitude > 0.05).all()
    assert np.abs(logs[False]['joints'][..., 0]).max() > 0.02          # the swimmer moves
    from farms_mujoco_b200.layout import sc
    assert np.abs(logs[False]['joints'][:, :, 5:, sc.joint_torque]).max() == 0   # motors are not logged (D-4)
    for kind in ('links', 'joints', 'xfrc'):
        err = log_error(kind, logs[True][kind], logs[False][kind])
        assert err < tol, (kind, err)
    # model.qpos_spring ends where the host would have left it, and it moved
    assert np.abs(springs[False][:, 7 + 5:]).max() > 0.05
    assert np.abs(springs[True] - springs[False]).max() < 1e-5, np.abs(springs[True] - springs[False]).max()


LEAN_TOL = 1e-5


def check_lean_variant(library, names=('swimmer8', 'salamander_swim', 'salamander'), n_envs=5, slims=(0,)):
    """LEAN variants of the unconstrained kernel (the model's unused paths compiled out) against the
    general one, in every layout: the same arithmetic -- the two are separate compilations, so the
    compiler may contract a multiply-add in one and not in the other: agreement to a few ulp
    (LEAN_TOL = 1e-5, group-wise relative, over 8 steps: drag forces square the velocity differences); the LAYOUTS of the general variant give the same
    bits, those of the LEAN variant agree to the same tolerance.
    Models outside the subset report fast_lean = False."""
    import variant_models
    from farms_mujoco_b200.engine import BatchedPhysics
    for name in names:
        spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs)
        outs = {}
        for slim in slims:
            for lean in (True, False):
                physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=9, library=library)
                if slim:
                    physics.set_fast_slim(slim)
                physics.set_fast_lean(lean)
                assert physics.fast_lean == lean, name
                # one constrained kernel in every layout (its SPLIT variant exists beside the regular
                # layout only and agrees to rounding, check_con_split_variant)
                physics.set_con_split(False)
                physics.reset(qpos0, qvel0)
                physics.set_ctrl(ctrl)
                physics.step(5)
                physics.step(3)
                outs[slim, lean] = (physics.qpos, physics.qvel, physics.xfrc_applied, physics.log_arrays())
        for slim in slims:
            a, b = outs[slim, True], outs[slim, False]
            for env in range(n_envs):
                errs = state_errors(a[0][env], a[1][env], b[0][env], b[1][env])
                assert max(errs.values()) < LEAN_TOL, (name, slim, errs)
                for kind in ('links', 'joints', 'contacts', 'xfrc'):
                    # contact forces: the solver may stop an iteration apart in the two compilations
                    tol = 5e-4 if kind == 'contacts' else LEAN_TOL
                    assert log_error(kind, a[3][kind][env], b[3][kind][env]) < tol, (name, slim, kind)
            # layouts of the GENERAL variant: same bits (one source, same contraction); of the LEAN
            # variant: the same to a few ulp
            ref, other = outs[slims[0], False], outs[slim, False]
            assert np.array_equal(ref[0], other[0]) and np.array_equal(ref[1], other[1]), (name, slim)
            assert np.array_equal(ref[2], other[2])
            for kind in ('links', 'joints', 'contacts', 'xfrc'):
                assert np.array_equal(ref[3][kind], other[3][kind]), (name, slim, kind)
            ref, other = outs[slims[0], True], outs[slim, True]
            for env in range(n_envs):
                assert max(state_errors(other[0][env], other[1][env], ref[0][env], ref[1][env]).values()) < LEAN_TOL
                for kind in ('links', 'joints', 'contacts', 'xfrc'):
                    tol = 5e-4 if kind == 'contacts' else LEAN_TOL
                    assert log_error(kind, other[3][kind][env], ref[3][kind][env]) < tol, (name, slim, kind)
    physics = BatchedPhysics.from_spec(variant_models.swimmer8_features(), 2, buffer_size=2, library=library)
    assert not physics.fast_lean            # slide joint, clamps, off-origin anchors


def check_split_variant_is_bit_identical(library, n_envs=75):
    """SPLIT variant (several warps per 32 environments, each its own bodies of the tree) against the
    single-warp kernel: the same per-body code in the same order along every chain, hence the same
    bits -- unconstrained rollouts, ground models (handed over at step 0) and hand-overs in the
    middle of a launch."""
    from farms_mujoco_b200.engine import BatchedPhysics
    for name in ('salamander_swim', 'salamander', 'centipede'):
        if name == 'salamander_swim':
            # some environments reach a joint limit within the launch: mixed warps
            spec, model, qpos0, qvel0, ctrl = hand_over_case(name, n_envs)
        else:
            spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs)
        outs = []
        for split in (False, True):
            physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=12, library=library)
            physics.set_fast_split(split)
            assert bool(physics.fast_split) == split and (not split or physics.fast_split >= 3), name
            physics.reset(qpos0, qvel0)
            physics.set_ctrl(ctrl)
            physics.step(6)
            pending = physics.last_pending
            physics.step(5)
            outs.append((physics.qpos, physics.qvel, physics.xfrc_applied, physics.log_arrays(), pending, physics.flags))
        assert outs[0][4] == outs[1][4], (name, outs[0][4], outs[1][4])
        if name == 'salamander_swim':
            assert 0 < outs[0][4] < n_envs
        assert np.array_equal(outs[0][5], outs[1][5])
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), name
        assert np.array_equal(outs[0][2], outs[1][2]), name
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            assert np.array_equal(outs[0][3][kind], outs[1][3][kind]), (name, kind)


def check_con_split_variant(library, n_envs=75, n_steps=(6, 5), names=('salamander', 'centipede'), tol=None):
    """SPLIT variant of the constrained per-thread kernel (several warps per group of environments,
    each its own bodies; fb_fastc_split_kernel) against the single-warp kernel on walking batches:
    every environment is handed over before its first step, so every group is taken by the SPLIT
    variant (the last one is partial).  The line-search sums are added over the warps in another
    order than on one warp: agreement to rounding (the tolerance of the LEAN check on the salamander;
    the centipede, whose fp32 floor against the oracle is 3e-5, to 1e-4), same flags."""
    from farms_mujoco_b200.engine import BatchedPhysics
    tol = tol or {'salamander': LEAN_TOL, 'centipede': 1e-4}
    for name in names:
        spec, model, qpos0, qvel0, ctrl = make_case(name, n_envs)
        outs = []
        for split in (False, True):
            physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=sum(n_steps) + 1, library=library)
            physics.set_con_split(split)
            assert physics.con_split == split, name
            physics.reset(qpos0, qvel0)
            physics.set_ctrl(ctrl)
            pending = []
            for n in n_steps:
                physics.step(n)
                pending.append(physics.last_pending)
            outs.append((physics.qpos, physics.qvel, physics.xfrc_applied, physics.log_arrays(), pending, physics.flags))
        assert outs[0][4] == outs[1][4] == [n_envs]*len(n_steps), (name, outs[0][4], outs[1][4])
        assert np.array_equal(outs[0][5], outs[1][5]), name
        a, b = outs
        worst = {}
        for env in range(n_envs):
            errs = state_errors(a[0][env], a[1][env], b[0][env], b[1][env])
            for kind in ('links', 'joints', 'contacts', 'xfrc'):
                for group, val in log_errors(kind, a[3][kind][env], b[3][kind][env]).items():
                    errs[kind + '.' + group] = val
            for key, val in errs.items():
                worst[key] = max(worst.get(key, 0.0), val)
        for key, val in worst.items():
            # constraint forces (contacts rows, joint limit force): the solver may stop an iteration
            # apart in the two variants, as in the LEAN check
            force = key.startswith('contacts.') or key == 'joints.limit_force'
            assert val < (5e-4 if force else tol[name]), (name, key, worst)
        # the contacts rows are not all zero: the comparison above is about real forces
        assert np.abs(a[3]['contacts']).max() > 0, name


def check_con_split_mixed_groups(library, n_envs=64, lift=0.06, fall=-1.0, n_steps=(1, 6, 6, 6), tol=5e-5):
    """The SPLIT constrained kernel takes only the groups whose environments were ALL handed over
    before their first step; the single-warp kernel behind it takes the others.  The first half of
    the batch stands on the ground; the second half starts `lift` above it (just outside the
    conservative plane bound of the hand-over test) moving down at `fall` m/s and crosses the bound
    in the middle of the third launch: in that launch every environment is handed over, the first
    half at step 0 (SPLIT), the second wherever it crossed (single warp).  Against the same batch
    with the SPLIT variant switched off: rounding-level differences of the two summation orders,
    grown over 19 steps of ground contact (the tolerance of the 20-step oracle comparisons)."""
    from farms_mujoco_b200.engine import BatchedPhysics
    spec, model, qpos0, qvel0, ctrl = make_case('salamander', n_envs, qvel_scale=0.05)
    qpos0[n_envs//2:, 2] += lift
    qvel0[n_envs//2:, 2] = fall
    outs, all_pending = [], []
    for split in (False, True):
        physics = BatchedPhysics.from_spec(spec, n_envs, buffer_size=sum(n_steps) + 1, library=library)
        physics.set_con_split(split)
        physics.reset(qpos0, qvel0)
        physics.set_ctrl(ctrl)
        pending = []
        for n in n_steps:
            physics.step(n)
            pending.append(physics.last_pending)
        all_pending.append(pending)
        outs.append((physics.qpos, physics.qvel, physics.log_arrays(), physics.flags))
    assert all_pending[0] == all_pending[1], all_pending
    # the falling half is outside the bound for the first launches and joins in the middle of one
    assert all_pending[0][0] == all_pending[0][1] == n_envs//2 and all_pending[0][2] == n_envs, all_pending
    assert np.array_equal(outs[0][3], outs[1][3])
    a, b = outs
    for env in range(n_envs):
        errs = state_errors(a[0][env], a[1][env], b[0][env], b[1][env])
        assert max(errs.values()) < tol, (env, errs)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            for group, val in log_errors(kind, a[2][kind][env], b[2][kind][env]).items():
                force = kind == 'contacts' or group == 'limit_force'
                assert val < (5e-4 if force else tol), (env, kind, group, val)
    assert np.abs(a[2]['contacts'][:n_envs//2]).max() > 0
