"""The reference-facing Python layer (Simulation / ExperimentTask / TaskCallback /
SwimmingHandler / physics2data mirrors) on the host-emulation build: hook order,
iteration bookkeeping, per-iteration vs fused stepping, parity with the oracle's replay of
``Simulation.run`` (oracle/farms_oracle.py: reference_rollout)."""

import os

import numpy as np

from conftest import scaled_error, log_error
from farms_mujoco_b200 import models, mjcf_subset
from farms_mujoco_b200.control import AnimatController, ControlType, TravellingWaveController
from farms_mujoco_b200.models import travelling_wave_parameters
from farms_mujoco_b200.simulation.simulation import Simulation
from farms_mujoco_b200.simulation.task import ExperimentTask, TaskCallback
from farms_mujoco_b200.swimming.drag import SwimmingHandler, WaterProperties
from farms_mujoco_b200.sensors.sensors import cycontacts2data


class Recorder(TaskCallback):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.calls = []

    def initialize_episode(self, task, physics):
        self.calls.append(('init', task.iteration))

    def before_step(self, task, action, physics):
        self.calls.append(('before', task.iteration))

    def after_step(self, task, physics):
        self.calls.append(('after', task.iteration))


class SwimmingCallback(TaskCallback):
    """The swimming callback a farms experiment attaches (farms_sim): a ``SwimmingHandler`` stepped
    in ``before_step``, on sub-steps too."""

    def __init__(self, spec):
        super().__init__(substep=True)
        self.spec, self.handler = spec, None

    def initialize_episode(self, task, physics):
        self.handler = SwimmingHandler(task.data, self.spec.animat_options, self.spec.arena_options,
                                       self.spec.simulation_options.units, physics)

    def before_step(self, task, action, physics):
        self.handler.step(task.iteration % task.buffer_size)


class HostWave(AnimatController):
    """A controller with no device form: evaluated on the host every iteration."""

    def __init__(self, joints, amp, freq, lag, phase):
        super().__init__(joints_names=[list(joints), [], []])
        self.amp, self.freq, self.lag, self.phase = amp, freq, lag, np.asarray(phase)

    def positions(self, iteration, time, timestep):
        values = self.amp*np.sin(2*np.pi*self.freq*time - self.lag + self.phase[:, None])
        return {j: values[:, i] for i, j in enumerate(self.joints_names[ControlType.POSITION])}


def test_hook_order_and_iteration_bookkeeping(emu_library):
    """task.py:168-186,348-369 + Appendix B: the first step resets; n calls = n-1 steps."""
    spec = models.swimmer8(n_iterations=6)
    cb = Recorder()
    sim = Simulation.from_spec(spec, n_envs=2, callbacks=[cb], library=emu_library)
    sim.run()
    assert cb.calls[0] == ('init', 0)
    assert cb.calls[1:5] == [('before', 0), ('after', 1), ('before', 1), ('after', 2)]
    assert len(cb.calls) == 1 + 2*5 and sim.iteration == 5
    assert sim.task.sim_iterations == 6 and sim.physics.iteration == 5
    assert isinstance(sim.task, ExperimentTask)
    assert sim.task.data.sensors.links.array.shape == (2, 6, 8, 20)
    assert sim.task.data.sensors.links.array.dtype == np.float64
    assert np.allclose(sim.task.data.sensors.links.masses, 0.15917403, atol=1e-6)


def test_per_iteration_loop_matches_reference_replay(emu_library):
    """Host controller, one launch per iteration: the reference's ordering exactly."""
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    n_it, phase = 10, [0.3, 1.1]
    spec = models.salamander(swimming=True, n_iterations=n_it)
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    sim = Simulation.from_spec(spec, n_envs=2, controller=HostWave(joints, amp, freq, lag, phase),
                               library=emu_library)
    assert sim.task.device_controller is False
    sim.run()
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    for env in range(2):
        def controller(iteration, time, env=env):
            ctrl = np.zeros(model.nu)
            ctrl[acts] = amp*np.sin(2*np.pi*freq*time - lag + phase[env])
            return ctrl
        data, states = fo.reference_rollout(OraclePhysics(model), spec, sim.physics.tables, n_it,
                                            controller=controller)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            ours = getattr(sim.task.data.sensors, kind).array[env]
            assert log_error(kind, ours, getattr(data.sensors, kind).array) < 2e-5, kind
        assert scaled_error(sim.physics.qpos[env], states[-1][0]) < 2e-5


def test_fused_launches_equal_per_iteration_loop(emu_library):
    spec = models.swimmer8(n_iterations=40)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    logs = []
    for callbacks, chunk in (([Recorder()], 0), ([], 16), ([], 1)):
        ctl = TravellingWaveController(joints, amp, freq, lag, env_phase=[0.0, 0.5, 1.0])
        sim = Simulation.from_spec(spec, n_envs=3, controller=ctl, callbacks=callbacks, chunk=chunk,
                                   library=emu_library)
        sim.run()
        assert sim.task.device_controller and sim.iteration == 39
        logs.append({k: getattr(sim.task.data.sensors, k).array.copy() for k in ('links', 'joints', 'xfrc')})
    for other in logs[1:]:
        for kind, arr in logs[0].items():
            assert np.array_equal(arr, other[kind]), kind


def test_iterator_and_postprocess(emu_library, tmp_path):
    spec = models.swimmer8(n_iterations=5)
    sim = Simulation.from_spec(spec, n_envs=1, library=emu_library)
    seen = list(sim.iterator(show_progress=False))
    assert seen == [0, 1, 2, 3, 4] and sim.iteration == 4
    times = sim.postprocess(sim.iteration, log_path=str(tmp_path))
    assert len(times) == 4
    saved = np.load(os.path.join(tmp_path, 'simulation.npz'))
    assert saved['links'].shape == (1, 4, 8, 20) and list(saved['links_names']) == spec.links_names
    assert os.path.exists(os.path.join(tmp_path, 'simulation_options.yaml'))
    assert os.path.exists(os.path.join(tmp_path, 'animat_options.yaml'))


def test_swimming_handler_and_contact_mirrors(emu_library):
    """drag.pyx:309-419 / sensors.pyx:140-190 protocol on top of the fused device step."""
    spec = models.salamander(swimming=True, n_iterations=4)
    sim = Simulation.from_spec(spec, n_envs=2, library=emu_library)
    sim.step()                                   # reset
    task, physics = sim.task, sim.physics
    handler = SwimmingHandler(task.data, spec.animat_options, spec.arena_options,
                              spec.simulation_options.units, physics)
    assert handler.n_links == len(spec.xfrc_names) and handler.drag and handler.buoyancy
    assert isinstance(handler.water, WaterProperties) and handler.water.surface() == 0.0
    sim.step()
    handler.step(1)
    assert task.data.sensors.xfrc.array[:, 1].any()
    assert np.array_equal(task.data.sensors.xfrc.array[:, 1], physics.log_row('xfrc', 1).astype(np.float64))
    handler.set_water_velocity([0.2, 0.0, 0.0])
    assert handler.water.velocity().tolist() == [0.2, 0.0, 0.0]
    cycontacts2data(physics, 1, task.data.sensors.contacts, task.maps['sensors']['geompair2data'], 1.0, 1.0)
    assert not task.data.sensors.contacts.array[:, 1].any()      # swimming: no contact


def test_physics_error_is_raised_or_handled(emu_library):
    from farms_mujoco_b200.simulation.simulation import PhysicsError
    import pytest
    spec = models.swimmer8(n_iterations=4)
    bad = np.tile(mjcf_subset.parse_mjcf(spec.mjcf).key_qvel, (1, 1)).astype(float)
    bad[0, 0] = np.inf
    sim = Simulation.from_spec(spec, n_envs=1, qvel0=bad, library=emu_library)
    with pytest.raises(PhysicsError):
        sim.run()
    sim = Simulation.from_spec(spec, n_envs=1, qvel0=bad, handle_exceptions=True, library=emu_library)
    sim.run()     # swallowed, as simulation.py:157-161 does with handle_exceptions


class HostTorque(AnimatController):
    """Torque control + spring references, evaluated on the host (task.py:323-346)."""

    def __init__(self, joints, n_envs):
        super().__init__(joints_names=[[], [], list(joints)])
        self.n_envs = n_envs

    def torques(self, iteration, time, timestep):
        return {j: 0.002*np.sin(40*time + i + np.arange(self.n_envs))
                for i, j in enumerate(self.joints_names[ControlType.TORQUE])}

    def springrefs(self, iteration, time, timestep):
        return {self.joints_names[ControlType.TORQUE][0]: 0.1}


def _torque_control_case(library):
    """initialize_control (task.py:262-286): motors without the 'position' control type get
    their position / velocity actuators force-limited to [0, 0]; torques reach ctrl scaled by
    units.torques; springrefs land in qpos_spring."""
    import copy
    import dataclasses
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    n_it, n_envs = 8, 2
    spec = models.swimmer8(n_iterations=n_it)
    animat = copy.deepcopy(spec.animat_options)
    for motor in animat.control.motors:
        motor.control_types = ['torque']
    spec = dataclasses.replace(spec, animat_options=animat)
    sim = Simulation.from_spec(spec, n_envs=n_envs, controller=HostTorque(spec.joints_names, n_envs),
                               library=library)
    sim.run()
    model = sim.physics.model
    pos = [model.actuator_id(f'actuator_position_{j}') for j in spec.joints_names]
    vel = [model.actuator_id(f'actuator_velocity_{j}') for j in spec.joints_names]
    trq = [model.actuator_id(f'actuator_torque_{j}') for j in spec.joints_names]
    assert np.asarray(model.actuator_forcelimited)[pos + vel].all()
    assert not np.asarray(model.actuator_forcerange).reshape(-1, 2)[pos + vel].any()
    assert not np.asarray(model.actuator_forcelimited)[trq].any()
    assert sim.physics.qpos_spring[:, 7].tolist() == [np.float32(0.1)]*n_envs
    controller = HostTorque(spec.joints_names, n_envs)
    for env in range(n_envs):
        def ctrl_of(iteration, time, env=env):
            ctrl = np.zeros(model.nu)
            values = controller.torques(iteration, time, model.timestep)
            for act, joint in zip(trq, spec.joints_names):
                ctrl[act] = values[joint][env]
            return ctrl
        oracle = OraclePhysics(model)          # the edited model
        oracle.model.qpos_spring[7] = 0.1
        data, _ = fo.reference_rollout(oracle, spec, sim.physics.tables, n_it, controller=ctrl_of)
        for kind in ('links', 'joints', 'xfrc'):
            ours = getattr(sim.task.data.sensors, kind).array[env]
            assert log_error(kind, ours, getattr(data.sensors, kind).array) < 2e-5, kind


def test_torque_control_disables_position_actuators(emu_library):
    _torque_control_case(emu_library)


def test_open_loop_host_controller_is_fused(emu_library):
    """A host controller that declares ``open_loop`` runs in fused launches (its outputs travel as
    a ctrl sequence) and produces exactly the per-iteration loop's log."""
    spec = models.swimmer8(n_iterations=30)
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    logs = []
    for open_loop in (False, True):
        ctl = HostWave(joints, amp, freq, lag, [0.2, 0.9])
        ctl.open_loop = open_loop
        sim = Simulation.from_spec(spec, n_envs=2, controller=ctl, chunk=8, library=emu_library)
        sim.run()
        assert not sim.task.device_controller and sim.iteration == 29
        launches = sim.physics.launch_count()
        logs.append(({k: getattr(sim.task.data.sensors, k).array.copy() for k in ('links', 'joints', 'xfrc')},
                     launches))
    assert logs[1][1] < logs[0][1]            # fewer launches when fused
    for kind, arr in logs[0][0].items():
        assert np.array_equal(arr, logs[1][0][kind]), kind


def _substep_sim(spec, substeps, library, n_envs=2, **kwargs):
    import dataclasses
    opts = dataclasses.replace(spec.simulation_options, num_sub_steps=substeps,
                               timestep=spec.simulation_options.timestep*substeps)
    return Simulation(mjcf_model=spec.mjcf, base_link=spec.base_link, simulation_options=opts,
                      animat_options=spec.animat_options, arena_options=spec.arena_options, n_envs=n_envs,
                      links_names=spec.links_names, joints_names=spec.joints_names,
                      contacts_names=spec.contacts_names, xfrc_names=spec.xfrc_names, library=library, **kwargs)


def _substep_check(sim, spec, n_it, substeps, phase, wave, n_sub_steps=1, tol=2e-5):
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    joints, amp, freq, lag = wave
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    for env in range(sim.physics.n_envs):
        def controller(iteration, time, env=env):
            ctrl = np.zeros(model.nu)
            ctrl[acts] = amp*np.sin(2*np.pi*freq*time - lag + phase[env])
            return ctrl
        data, (ref_q, _) = fo.reference_rollout_substeps(
            OraclePhysics(model), spec, sim.physics.tables, n_it, substeps,
            timestep=sim.options.timestep, controller=controller, n_sub_steps=n_sub_steps)
        for kind in ('links', 'joints', 'contacts', 'xfrc'):
            ours = getattr(sim.task.data.sensors, kind).array[env]
            ref = getattr(data.sensors, kind).array
            assert ours.shape == ref.shape
            assert log_error(kind, ours, ref) < tol, kind
        assert scaled_error(sim.physics.qpos[env], ref_q) < tol


def test_sub_steps_log_full_steps_only(emu_library):
    """num_sub_steps = 2: the device ring holds a row per physics step, the host rows
    are the reference's -- the state at every full step, row 0 the reset state, control held
    between full steps, drag refreshed every sub-step -- against the oracle's literal replay of
    task.py:168-186, 348-369."""
    n_it, phase = 6, [0.3, 1.1]
    spec = models.salamander(swimming=True, n_iterations=n_it)
    wave = travelling_wave_parameters(spec)
    sim = _substep_sim(spec, 2, emu_library, controller=HostWave(*wave, phase))
    assert sim.physics.log_stride == 2 and sim.physics.device_ring == 2*n_it
    sim.run()
    assert sim.task.sim_iterations == 2*n_it and sim.physics.iteration == 2*n_it - 1
    assert sim.iteration == n_it
    _substep_check(sim, spec, n_it, 2, phase, wave)
    # the device-side accessors return the same rows
    logs = sim.physics.log_arrays()
    assert logs['links'].shape[1] == n_it
    assert np.array_equal(logs['joints'][:, 2], sim.task.data.sensors.joints.array[:, 2].astype(np.float32))
    exported = sim.physics.export_farms(1)
    assert np.array_equal(exported.sensors.links.array[3], logs['links'][1, 3].astype(np.float64))


def test_sub_steps_with_a_substep_callback(emu_library):
    """num_sub_steps = 3 and a callback with substep=True: the links (and the drag forces computed
    from them) are refreshed on every sub-step into the row the reference's bookkeeping points at
    -- the current row on the first sub-step, the next row on the last (task.py:358-360) -- and
    the device controller is replaced by the host's, evaluated on full steps."""
    n_it, phase = 5, [0.0, 0.7]
    spec = models.salamander(swimming=True, n_iterations=n_it)
    wave = travelling_wave_parameters(spec)
    ctl = TravellingWaveController(*wave, env_phase=phase)
    cb = Recorder(substep=True)
    sim = _substep_sim(spec, 3, emu_library, controller=ctl, callbacks=[cb, SwimmingCallback(spec)])
    sim.run()
    assert sim.task.device_controller is False and sim.task.substeps_links
    assert [c for c in cb.calls if c[0] == 'before'][:4] == [('before', 0), ('before', 0), ('before', 1), ('before', 1)]
    _substep_check(sim, spec, n_it, 3, phase, wave)
    links = sim.task.data.sensors.links.array
    joints = sim.task.data.sensors.joints.array
    # row 1: joints from the full step (3 physics steps), links from the sub-step after it (4)
    assert np.allclose(joints[:, 1, :, 0], sim.physics.log_row('joints', 1)[:, :, 0])
    assert not np.allclose(links[:, 1], sim.physics.log_row('links', 1))


def test_dm_control_n_sub_steps(emu_library):
    """n_sub_steps = 2 (dm_control's Environment: two physics steps inside one env.step) on the
    ground; with swimming links it is refused (the reference holds the drag forces across them)."""
    import pytest
    n_it, phase = 40, [0.2, 0.9]
    spec = models.salamander(n_iterations=n_it)
    wave = travelling_wave_parameters(spec)
    sim = Simulation.from_spec(spec, n_envs=2, controller=HostWave(*wave, phase), n_sub_steps=2,
                               library=emu_library)
    sim.run()
    assert sim.physics.iteration == 2*(n_it - 1)
    _substep_check(sim, spec, n_it, 1, phase, wave, n_sub_steps=2, tol=1e-4)
    assert sim.task.data.sensors.contacts.array.any()
    with pytest.raises(NotImplementedError):
        Simulation.from_spec(models.swimmer8(n_iterations=4), n_envs=1, n_sub_steps=2, library=emu_library)
