"""Task (mirror of farms_mujoco/simulation/task.py for the batched engine).

Same names, arguments and hook order as the reference's ``ExperimentTask`` and
``TaskCallback``; ``physics`` is a ``BatchedPhysics`` (n_envs environments in
lockstep) and ``data`` a ``BatchedAnimatData`` (the reference arrays with a
leading environment axis).  With ``n_envs=1`` and one physics step per call
the sequence of operations is the reference's exactly:

    before_step:  sensors -> callbacks -> control      (task.py:168-186)
    physics.step                                       (simulation.py:156)
    after_step:   counters -> completion -> callbacks  (task.py:348-369)

What moved to the device: ``physics2data`` (the log row is produced by the
step kernel; ``update_sensors`` copies it), the swimming callback (fused) and,
when the controller offers ``device_parameters()``, the controller itself.
Viewer, height fields and muscles are outside the path.
"""

import numpy as np

from ..control import ControlType
from ..data import BatchedAnimatData
from ..units import SimulationUnitScaling as SimulationUnits
from .physics import physics2data


def duration2nit(duration, timestep):
    """Number of iterations from duration (task.py:32-34)"""
    return int(duration/timestep)


class ExperimentTask:
    """FARMS experiment (task.py:37-412)"""
    # pylint: disable=too-many-instance-attributes

    def __init__(self, base_link, n_iterations, timestep, **kwargs):
        self._app = None
        self.iteration = 0
        self.timestep = timestep
        self.n_iterations = n_iterations
        self.base_link = base_link
        self.data = kwargs.pop('data', None)
        self._controller = kwargs.pop('controller', None)
        self.animat_options = kwargs.pop('animat_options', None)
        self.external_force = kwargs.pop('external_force', 0.2)
        self._restart = kwargs.pop('restart', True)
        self._callbacks = kwargs.pop('callbacks', [])
        self._extras = {'hfield': kwargs.pop('hfield', None)}
        self.units = kwargs.pop('units', SimulationUnits())
        self.substeps = max(1, kwargs.pop('substeps', 1))
        self.buffer_size = max(1, kwargs.pop('buffer_size', 1))
        self.substeps_links = any(cb.substep for cb in self._callbacks)
        self.sim_iteration = 0
        self.sim_iterations = self.n_iterations*self.substeps
        self.sim_timestep = self.timestep/self.substeps
        self.maps = {
            'sensors': {}, 'ctrl': {},
            'xpos': {}, 'qpos': {}, 'geoms': {},
            'links': {}, 'joints': {}, 'contacts': {}, 'xfrc': {},
            'muscles': {},
        }
        self.device_controller = False      # the controller runs inside the step kernel
        # On-device controllers advance with every physics step; with several physics steps per
        # iteration the reference holds ctrl between full steps (task.py:184-186), so the host
        # evaluates the controller then (Simulation passes device_control=False).
        self._device_control = bool(kwargs.pop('device_control', True))
        assert not kwargs, kwargs
        assert self._extras['hfield'] is None, 'height fields are outside the batched path'

    @property
    def controller(self):
        return self._controller

    @property
    def callbacks(self):
        return self._callbacks

    def initialize_episode(self, physics):
        """Sets the state of the environment at the start of each episode (task.py:87-154)"""
        # Links masses
        self.initialize_maps(physics)
        if self.data is None:
            self.initialize_data(physics)
        model = physics.model
        self.data.sensors.links.masses = np.array([
            model.body_mass[model.body_id(link_name)]
            for link_name in self.data.sensors.links.names
        ], dtype=float)/self.units.kilograms
        # Initialise iterations
        self.iteration = 0
        self.sim_iteration = 0
        # Maps, data and sensors
        self.initialize_sensors(physics)
        # Control
        if self._controller is not None:
            self.initialize_control(physics)
        # Initialize joints to keyframe 0 (+ the reset's mj_forward: log row 0)
        physics.reset()
        # Callbacks
        for callback in self._callbacks:
            callback.initialize_episode(task=self, physics=physics)

    def update_sensors(self, physics, links_only=False):
        """Update sensors (task.py:156-166)"""
        index = self.iteration % self.buffer_size
        physics2data(physics=physics, iteration=index, data=self.data, maps=self.maps,
                     units=self.units, links_only=links_only)

    def before_step(self, action, physics):
        """Operations before physics step (task.py:168-186)"""
        assert self.iteration < self.n_iterations
        # Sensors
        full_step = not self.sim_iteration % self.substeps
        if full_step or self.substeps_links:
            self.update_sensors(physics=physics, links_only=not full_step)
        # Callbacks
        for callback in self._callbacks:
            if full_step or callback.substep:
                callback.before_step(task=self, action=action, physics=physics)
        # Control
        if full_step and self._controller is not None and not self.device_controller:
            self.step_control(physics)

    def initialize_maps(self, physics):
        """Initialise data (task.py:188-206)"""
        model = physics.model
        self.maps['xpos']['names'] = list(model.body_names)
        self.maps['qpos']['names'] = list(model.jnt_names)
        self.maps['xfrc']['names'] = list(model.body_names)
        self.maps['geoms']['names'] = list(model.geom_names)
        self.maps['muscles']['names'] = []

    def initialize_data(self, physics):
        """Initialise data (task.py:208-218): the engine's own sensor names"""
        names = physics.names
        self.data = BatchedAnimatData(
            timestep=self.timestep, n_envs=physics.n_envs, buffer_size=self.buffer_size,
            links=names.links.names, joints=names.joints.names, contacts=names.contacts.names,
            xfrc=names.xfrc.names, dtype=np.float64)

    def initialize_sensors(self, physics):
        """Initialise sensors (task.py:218-225): the maps the engine was compiled with"""
        self.maps['sensors'] = physics.maps['sensors']

    def initialize_control(self, physics):
        """Initialise controller (task.py:227-286)"""
        model = physics.model
        ctrl_names = list(model.actuator_names)
        for joint in self._controller.joints_names[ControlType.POSITION]:
            assert f'actuator_position_{joint}' in ctrl_names, f'{joint} not in {ctrl_names}'
        self.maps['ctrl']['pos'] = [
            ctrl_names.index(f'actuator_position_{joint}')
            for joint in self._controller.joints_names[ControlType.POSITION]]
        self.maps['ctrl']['vel'] = [
            ctrl_names.index(f'actuator_velocity_{joint}')
            for joint in self._controller.joints_names[ControlType.VELOCITY]]
        self.maps['ctrl']['trq'] = [
            ctrl_names.index(f'actuator_torque_{joint}')
            for joint in self._controller.joints_names[ControlType.TORQUE]]
        assert not self._controller.muscles_names, 'muscles are outside the batched path'
        self.maps['ctrl']['springref'] = {
            joint: int(model.jnt_qposadr[model.jnt_id(joint)])
            for joint in model.jnt_names if model.jnt_type[model.jnt_id(joint)] != 0}
        # Actuator limits (task.py:252-286): position / velocity actuators of joints whose motor
        # has no 'position' control type are force-limited to [0, 0] -- a model edit
        jntname2actid = {name: {} for name in model.jnt_names}
        biasprm = np.asarray(model.actuator_biasprm).reshape(model.nu, -1)
        trnid = np.asarray(model.actuator_trnid).reshape(model.nu, -1)[:, 0]
        for act_i in range(model.nu):
            act_type = 'pos' if biasprm[act_i, 1] != 0 else 'vel' if biasprm[act_i, 2] != 0 else 'trq'
            jntname2actid[model.jnt_names[trnid[act_i]]][act_type] = act_i
        if self.animat_options is not None:
            disabled = []
            for mtr_opts in self.animat_options.control.motors:
                jnt_name = mtr_opts['joint_name']
                if 'position' not in mtr_opts.control_types:
                    disabled += [jntname2actid[jnt_name][t] for t in ('pos', 'vel')
                                 if t in jntname2actid[jnt_name]]
            if disabled:
                physics.set_actuator_forcerange(disabled, True, [0.0, 0.0])
        if self._device_control and hasattr(self._controller, 'device_parameters') and not self._callbacks_need_ctrl():
            params = self._controller.device_parameters()
            acts = [ctrl_names.index(f'actuator_position_{j}') for j in params['joints']]
            phase = np.broadcast_to(np.asarray(params['env_phase'], dtype=float), (physics.n_envs,))
            physics.set_env_phase(np.ascontiguousarray(phase))
            physics.set_wave_controller(acts, params['amplitude'], params['frequency'],
                                        params['phase_lag'], params.get('offset'))
            self.device_controller = True
        if self._device_control and callable(getattr(self._controller, 'device_cpg', None)) and not self._callbacks_need_ctrl():
            # coupled-oscillator network integrated on the device (fb_set_cpg): position targets and
            # torque commands (x units.torques, task.py:332) written inside every launch
            names = {ControlType.POSITION: 'actuator_position_{}', ControlType.TORQUE: 'actuator_torque_{}'}
            net = self._controller.device_cpg(
                lambda joint, kind: ctrl_names.index(names[ControlType(kind)].format(joint)),
                torque_unit=self.units.torques,
                spring_index=lambda joint: self.maps['ctrl']['springref'][joint])
            physics.set_cpg(net)
            physics.set_cpg_state(net['phase0'], net['amplitude0'])
            self.device_controller = True
        self._ctrl = np.zeros((physics.n_envs, model.nu))

    def _callbacks_need_ctrl(self):
        return any(getattr(cb, 'reads_ctrl', False) for cb in self._callbacks)

    def step_control(self, physics):
        """Step control (task.py:288-307)"""
        current_time = self.iteration*self.timestep
        index = self.iteration % self.buffer_size
        self._controller.step(iteration=index, time=current_time, timestep=self.timestep)
        if self._controller.joints_names[ControlType.POSITION]:
            self.step_joints_control_position(physics, current_time)
        if self._controller.joints_names[ControlType.TORQUE]:
            self.step_joints_control_torque(physics, current_time)
        physics.set_ctrl(self._ctrl)

    def control_sequence(self, n_steps):
        """``ctrl`` of the next ``n_steps`` iterations, ``[n_steps, n_envs, nu]``, for a controller
        whose output depends on (iteration, time) only (``controller.open_loop``): what
        ``step_control`` would write before each of those steps (task.py:288-346), computed
        ahead so that the engine can fuse the steps into one launch (fb_set_ctrl_sequence)."""
        sequence = np.empty((n_steps,) + self._ctrl.shape, dtype=np.float32)
        for k in range(n_steps):
            iteration = self.iteration + k
            current_time = iteration*self.timestep
            index = iteration % self.buffer_size
            self._controller.step(iteration=index, time=current_time, timestep=self.timestep)
            if self._controller.joints_names[ControlType.POSITION]:
                positions = self._controller.positions(iteration=index, time=current_time, timestep=self.timestep)
                for act, joint in zip(self.maps['ctrl']['pos'], self._controller.joints_names[ControlType.POSITION]):
                    self._ctrl[:, act] = positions[joint]
            if self._controller.joints_names[ControlType.TORQUE]:
                torques = self._controller.torques(iteration=index, time=current_time, timestep=self.timestep)
                for act, joint in zip(self.maps['ctrl']['trq'], self._controller.joints_names[ControlType.TORQUE]):
                    self._ctrl[:, act] = np.asarray(torques[joint])*self.units.torques
                assert not self._controller.springrefs(iteration=index, time=current_time, timestep=self.timestep), (
                    'spring references change qpos_spring: not expressible as a ctrl sequence')
            sequence[k] = self._ctrl
        return sequence

    def step_joints_control_position(self, physics, time):
        """Step position control (task.py:309-321)"""
        del physics
        index = self.iteration % self.buffer_size
        joints_positions = self._controller.positions(iteration=index, time=time, timestep=self.timestep)
        for act, joint in zip(self.maps['ctrl']['pos'], self._controller.joints_names[ControlType.POSITION]):
            self._ctrl[:, act] = joints_positions[joint]

    def step_joints_control_torque(self, physics, time):
        """Step torque control (task.py:323-346)"""
        index = self.iteration % self.buffer_size
        joints_torques = self._controller.torques(iteration=index, time=time, timestep=self.timestep)
        torques = self.units.torques
        for act, joint in zip(self.maps['ctrl']['trq'], self._controller.joints_names[ControlType.TORQUE]):
            self._ctrl[:, act] = np.asarray(joints_torques[joint])*torques
        springrefs = self._controller.springrefs(iteration=index, time=time, timestep=self.timestep)
        if springrefs:
            qpos_spring = physics.qpos_spring
            for joint, value in springrefs.items():
                qpos_spring[:, self.maps['ctrl']['springref'][joint]] = value
            physics.set_qpos_spring(qpos_spring)

    def after_step(self, physics):
        """Operations after physics step (task.py:348-369)"""
        self.sim_iteration += 1
        fullstep = not (self.sim_iteration + 1) % self.substeps
        if fullstep:
            self.iteration += 1
        assert self.iteration <= self.n_iterations
        if fullstep:
            for callback in self._callbacks:
                callback.after_step(task=self, physics=physics)

    def action_spec(self, physics):
        """Action specifications (task.py:371-378)"""
        specs = []
        for callback in self._callbacks:
            spec = callback.action_spec(task=self, physics=physics)
            if spec is not None:
                specs += spec
        return specs

    def step_spec(self, physics):
        """Timestep specifications"""
        for callback in self._callbacks:
            callback.step_spec(task=self, physics=physics)

    def get_observation(self, physics):
        """Environment observation"""
        for callback in self._callbacks:
            callback.get_observation(task=self, physics=physics)

    def get_reward(self, physics):
        """Reward (task.py:390-397)"""
        reward = 0
        for callback in self._callbacks:
            callback_reward = callback.get_reward(task=self, physics=physics)
            if callback_reward is not None:
                reward += callback_reward
        return reward

    def get_termination(self, physics):
        """Return final discount if episode should end, else None (task.py:399-407)"""
        terminate = None
        for callback in self._callbacks:
            if callback.get_termination(task=self, physics=physics):
                terminate = 1
        if self.iteration >= self.n_iterations:
            terminate = 1
        return terminate

    def observation_spec(self, physics):
        """Observation specifications"""
        for callback in self._callbacks:
            callback.observation_spec(task=self, physics=physics)


class TaskCallback:
    """Task callback (task.py:415-446): the nine hooks and the ``substep`` flag"""

    def __init__(self, substep=False):
        self.substep = substep

    def initialize_episode(self, task, physics):
        """Initialize episode"""

    def before_step(self, task, action, physics):
        """Before step"""

    def after_step(self, task, physics):
        """After step"""

    def action_spec(self, task, physics):
        """Action specifications"""

    def step_spec(self, task, physics):
        """Timestep specifications"""

    def get_observation(self, task, physics):
        """Environment observation"""

    def get_reward(self, task, physics):
        """Reward"""

    def get_termination(self, task, physics):
        """Return final discount if episode should end, else None"""

    def observation_spec(self, task, physics):
        """Observation specifications"""
