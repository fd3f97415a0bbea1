"""fp64 NumPy restatement of the farms_mujoco per-step data plane.

TEST INFRASTRUCTURE ONLY (see oracle/mjstep_oracle.c).  Each function cites the
reference lines it follows; these parts of the path ARE fully specified by the
reference source, unlike the MuJoCo part.

  physics2data          farms_mujoco/simulation/physics.py:527-545
  cycontacts2data       farms_mujoco/sensors/sensors.pyx:140-190
  drag_forces           farms_mujoco/swimming/drag.pyx:152-268
  SwimmingHandlerOracle farms_mujoco/swimming/drag.pyx:309-411
  apply_xfrc            downstream swimming callback
  reference_rollout     simulation.py:149-161 + task.py:168-186,288-369 + dm_control
                        Environment.step semantics
"""

import numpy as np

from farms_mujoco_b200.layout import sc
from farms_mujoco_b200.data import AnimatData
from farms_mujoco_b200.simulation.physics import get_sensor_maps, get_physics2data_maps
from farms_mujoco_b200.units import SimulationUnitScaling


# --------------------------------------------------------------------------
# farms_core.utils.transform restatements (xyzw quaternions, Appendix C)
# --------------------------------------------------------------------------

def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_mult(q0, q1):
    """Hamilton product in xyzw."""
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return np.array([
        w0*x1 + x0*w1 + y0*z1 - z0*y1,
        w0*y1 - x0*z1 + y0*w1 + z0*x1,
        w0*z1 + x0*y1 - y0*x1 + z0*w1,
        w0*w1 - x0*x1 - y0*y1 - z0*z1,
    ])


def quat_rot(vector, quat):
    """out = q (x) (v, 0) (x) q*  (drag.pyx:50-56 call signature)."""
    v4 = np.array([vector[0], vector[1], vector[2], 0.0])
    return quat_mult(quat_mult(quat, v4), quat_conj(quat))[:3]


# --------------------------------------------------------------------------
# physics -> data (physics.py:423-545)
# --------------------------------------------------------------------------

def physicslinks2data(physics, iteration, data, sensor_maps, units):
    """physics.py:449-466 -- note the CoM orientation is the *body* xquat."""
    links = data.sensors.links.array
    links[iteration, :, sc.link_urdf_position_x:sc.link_urdf_position_z+1] = (
        physics.data.xpos[sensor_maps['xpos2data']]/units.meters)
    links[iteration, :, sc.link_urdf_orientation_x:sc.link_urdf_orientation_w+1] = (
        physics.data.xquat[sensor_maps['xquat2data']][:, [1, 2, 3, 0]])
    links[iteration, :, sc.link_com_position_x:sc.link_com_position_z+1] = (
        physics.data.xipos[sensor_maps['xipos2data']]/units.meters)
    links[iteration, :, sc.link_com_orientation_x:sc.link_com_orientation_w+1] = (
        physics.data.xquat[sensor_maps['xquat2data']][:, [1, 2, 3, 0]])


def physicslinksvelsensors2data(physics, iteration, data, sensor_maps, units):
    """physics.py:435-446"""
    links = data.sensors.links.array
    links[iteration, :, sc.link_com_velocity_lin_x:sc.link_com_velocity_lin_z+1] = (
        physics.data.sensordata[sensor_maps['framelinvel2data']]/units.velocity)
    links[iteration, :, sc.link_com_velocity_ang_x:sc.link_com_velocity_ang_z+1] = (
        physics.data.sensordata[sensor_maps['frameangvel2data']]/units.angular_velocity)


def physicsjointssensors2data(physics, iteration, data, sensor_maps, units):
    """physics.py:481-497"""
    joints = data.sensors.joints.array
    if len(sensor_maps['jointlimitfrc2data']) > 0:
        joints[iteration, :, sc.joint_limit_force] = (
            physics.data.sensordata[sensor_maps['jointlimitfrc2data']]/units.torques)
    if len(sensor_maps['force2data']) > 0:
        joints[iteration, :, sc.joint_force_x:sc.joint_force_z+1] = (
            physics.data.sensordata[sensor_maps['force2data']]/units.newtons)
    if len(sensor_maps['torque2data']) > 0:
        joints[iteration, :, sc.joint_torque_x:sc.joint_torque_z+1] = (
            physics.data.sensordata[sensor_maps['torque2data']]/units.torques)


def physicsjoints2data(physics, iteration, data, sensor_maps, units):
    """physics.py:500-507"""
    joints = data.sensors.joints.array
    joints[iteration, :, sc.joint_position] = physics.data.qpos[sensor_maps['qpos2data']]
    joints[iteration, :, sc.joint_velocity] = (
        physics.data.qvel[sensor_maps['qvel2data']]/units.angular_velocity)


def physicsactuators2data(physics, iteration, data, sensor_maps, units):
    """physics.py:510-524 -- accumulating (+=); the torque map is empty (Appendix D-4)."""
    joints = data.sensors.joints.array
    itorques = 1.0/units.torques
    for key in ('actuatorfrc_position2data', 'actuatorfrc_velocity2data',
                'actuatorfrc_torque2data'):
        if len(sensor_maps[key]) > 0:
            joints[iteration, :, sc.joint_torque] += (
                physics.data.sensordata[sensor_maps[key]]*itorques)


def cycontacts2data(physics, iteration, data, geompair2data, meters, newtons):
    """sensors.pyx:140-190 (+ store_forces :20-52, postprocess_contacts :113-137)."""
    cdata = data.array
    norm_sum = np.zeros(cdata.shape[1])
    for contact_i, contact in enumerate(physics.data.contact):
        geom1, geom2 = contact.geom1, contact.geom2
        for pair, sign in (((geom1, geom2), -1), ((geom2, geom1), +1),
                           ((geom1, -1), -1), ((geom2, -1), +1)):
            if pair not in geompair2data:
                continue
            index = geompair2data[pair]
            forcetorque = physics.contact_force(contact_i)     # mj_contactForce, :70
            frame, pos = contact.frame, contact.pos
            reaction = sign*forcetorque[0]*frame[0:3]
            friction = sign*forcetorque[1]*frame[3:6] + sign*forcetorque[2]*frame[6:9]
            total = reaction + friction
            cdata[iteration, index, sc.contact_reaction_x:sc.contact_reaction_z+1] += reaction
            cdata[iteration, index, sc.contact_friction_x:sc.contact_friction_z+1] += friction
            cdata[iteration, index, sc.contact_total_x:sc.contact_total_z+1] += total
            norm = np.sqrt(total @ total)
            cdata[iteration, index, sc.contact_position_x:sc.contact_position_z+1] += norm*pos
            norm_sum[index] += norm
    for index in range(len(data.names)):
        if norm_sum[index] > 0:
            cdata[iteration, index, sc.contact_position_x:sc.contact_position_z+1] /= norm_sum[index]
        cdata[iteration, index, sc.contact_reaction_x:sc.contact_total_z+1] *= 1.0/newtons
        cdata[iteration, index, sc.contact_position_x:sc.contact_position_z+1] *= 1.0/meters


def physics2data(physics, iteration, data, maps, units, links_only=False):
    """physics.py:527-545"""
    sensor_maps = maps['sensors']
    physicslinks2data(physics, iteration, data, sensor_maps, units)
    physicslinksvelsensors2data(physics, iteration, data, sensor_maps, units)
    if not links_only:
        physicsjointssensors2data(physics, iteration, data, sensor_maps, units)
        physicsjoints2data(physics, iteration, data, sensor_maps, units)
        physicsactuators2data(physics, iteration, data, sensor_maps, units)
        cycontacts2data(physics, iteration, data.sensors.contacts,
                        sensor_maps['geompair2data'], units.meters, units.newtons)


# --------------------------------------------------------------------------
# swimming (drag.pyx)
# --------------------------------------------------------------------------

def drag_forces(iteration, links, links_index, xfrc, xfrc_index, coefficients, water,
                mass, height, density, gravity, use_buoyancy):
    """drag.pyx:152-268.  Returns False (row untouched) above the surface."""
    # pylint: disable=too-many-arguments,too-many-locals
    row = links[iteration, links_index]
    pos_z = row[2]
    surface = water['surface']
    if pos_z > surface:
        return False
    urdf2global = row[sc.link_urdf_orientation_x:sc.link_urdf_orientation_w+1]
    com2global = row[sc.link_com_orientation_x:sc.link_com_orientation_w+1]
    global2urdf = quat_conj(urdf2global)
    com2urdf = quat_mult(global2urdf, com2global)
    urdf2com = quat_conj(com2urdf)
    lin = quat_rot(row[sc.link_com_velocity_lin_x:sc.link_com_velocity_lin_z+1], global2urdf)
    ang = quat_rot(row[sc.link_com_velocity_ang_x:sc.link_com_velocity_ang_z+1], global2urdf)
    buoyancy = np.zeros(3)
    if use_buoyancy and mass > 0 and pos_z < surface:
        lift = -1000*mass*gravity/density*min(max(surface - pos_z, 0)/height, 1)
        buoyancy = quat_rot(np.array([0.0, 0.0, lift]), global2urdf)
    lin = lin - quat_rot(water['velocity'], global2urdf)
    force = np.sign(lin)*lin*lin*water['viscosity']*coefficients[0] + buoyancy
    # sign(0)*0 = 0 matches the branch "if v < 0: *= -1" at v == 0
    torque = np.sign(ang)*ang*ang*coefficients[1]
    xfrc[iteration, xfrc_index, 0:3] = quat_rot(force, urdf2com)
    xfrc[iteration, xfrc_index, 3:6] = quat_rot(torque, urdf2com)
    return True


class SwimmingHandlerOracle:
    """drag.pyx:309-411 driven by a ``FarmsTables`` (same constructor outputs)."""

    def __init__(self, data, tables):
        self.links = data.sensors.links
        self.xfrc = data.sensors.xfrc
        self.t = tables
        self.water = dict(surface=tables.water_surface, velocity=np.array(tables.water_velocity),
                          viscosity=tables.water_viscosity)

    def set_water_velocity(self, velocity):
        self.water['velocity'] = np.array(velocity, dtype=float)

    def step(self, iteration):
        t = self.t
        if not (t.water_drag or t.water_sph) or not t.water_drag:
            return
        for i in range(len(t.swim_links_index)):
            drag_forces(iteration, self.links.array, t.swim_links_index[i], self.xfrc.array,
                        t.swim_xfrc_index[i], t.swim_coefficients[i], self.water,
                        t.swim_mass[i], t.swim_height[i], t.swim_density[i],
                        gravity=-9.81, use_buoyancy=t.water_buoyancy)


def apply_xfrc(physics, data, iteration, sensor_maps, units):
    """Downstream swimming callback: link-local wrench -> world ``xfrc_applied``."""
    indices = sensor_maps['data2xfrc']
    physics.data.xfrc_applied[:, :] = 0
    for k, body in enumerate(indices):
        rot = physics.data.xmat[body].reshape(3, 3)
        wrench = data.sensors.xfrc.array[iteration, k]
        physics.data.xfrc_applied[body, 0:3] = rot @ wrench[0:3]*units.newtons
        physics.data.xfrc_applied[body, 3:6] = rot @ wrench[3:6]*units.torques


# --------------------------------------------------------------------------
# the reference loop
# --------------------------------------------------------------------------

def make_maps(model, data):
    sensor_maps = get_sensor_maps(model)
    get_physics2data_maps(model, data.sensors, sensor_maps)
    ctrl_names = list(model.actuator_names)
    return {'sensors': sensor_maps, 'ctrl_names': ctrl_names}


def reference_rollout(physics, spec, tables, n_iterations, controller=None, units=None,
                      swimming=True, qpos0=None, qvel0=None):
    """Replay ``Simulation.run`` for one environment on the CPU oracle.

    Order per ``env.step`` (task.py:168-186): sensors -> callbacks (swimming) ->
    control -> ``mj_step``; the first ``env.step`` is the reset (Appendix B), so
    ``n_iterations`` calls make ``n_iterations - 1`` physics steps and fill log
    rows ``0 .. n_iterations-1``.  ``controller(iteration, time)`` returns the
    full ``ctrl`` vector (or None).  Returns the ``AnimatData`` log.
    """
    # pylint: disable=too-many-arguments,too-many-locals
    units = units if units is not None else SimulationUnitScaling()
    model = physics.model
    data = AnimatData.from_sensors_names(
        timestep=model.timestep, buffer_size=n_iterations, links=spec.links_names,
        joints=spec.joints_names, contacts=spec.contacts_names, xfrc=spec.xfrc_names)
    maps = make_maps(model, data)
    handler = SwimmingHandlerOracle(data, tables)
    physics.reset(keyframe_id=0)
    if qpos0 is not None:
        physics.data.qpos[:] = qpos0
        if qvel0 is not None:
            physics.data.qvel[:] = qvel0
        physics.forward()
    states = [(physics.data.qpos.copy(), physics.data.qvel.copy())]
    for iteration in range(n_iterations):
        physics2data(physics, iteration, data, maps, units)
        if swimming and len(tables.swim_links_index):
            handler.step(iteration)
            apply_xfrc(physics, data, iteration, maps['sensors'], units)
        if controller is not None:
            ctrl = controller(iteration, iteration*model.timestep)
            if ctrl is not None:
                physics.data.ctrl[:] = ctrl
        if iteration == n_iterations - 1:
            # the reference stops one row earlier (its first env.step is the
            # reset); this extra row is what one more env.step would log
            break
        physics.step()
        states.append((physics.data.qpos.copy(), physics.data.qvel.copy()))
    return data, states


def reference_rollout_substeps(physics, spec, tables, n_iterations, substeps, timestep, controller=None,
                               units=None, swimming=True, n_sub_steps=1, qpos0=None, qvel0=None):
    """``Simulation.run`` with ``num_sub_steps = substeps`` physics steps per iteration, followed
    literally: ``before_step`` (task.py:168-186: full_step = not sim_iteration % substeps, the
    sensors refreshed on full steps and -- links only -- on sub-steps because the swimming callback
    has ``substep=True``, the callback every sub-step, control on full steps with time =
    iteration*timestep), ``n_sub_steps`` x ``mj_step``, ``after_step`` (task.py:348-369: iteration
    advances when ``(sim_iteration + 1) % substeps == 0``, i.e. BEFORE the last sub-step of an
    iteration, so that sub-step's links refresh lands in the next row and, for substeps >= 3, the
    earlier ones overwrite the links of the current row).  ``timestep`` is the iteration's
    (``options.timestep``); the model's is ``timestep/substeps`` (mjcf.py:1189).  The first
    ``env.step`` is the reset, so ``n_iterations*substeps - 1`` steps are taken; the row of the
    state after the last one is added when the buffer still has it."""
    # pylint: disable=too-many-arguments,too-many-locals
    units = units if units is not None else SimulationUnitScaling()
    model = physics.model
    data = AnimatData.from_sensors_names(
        timestep=timestep, buffer_size=n_iterations, links=spec.links_names,
        joints=spec.joints_names, contacts=spec.contacts_names, xfrc=spec.xfrc_names)
    maps = make_maps(model, data)
    handler = SwimmingHandlerOracle(data, tables)
    swim = swimming and len(tables.swim_links_index)
    physics.reset(keyframe_id=0)
    if qpos0 is not None:
        physics.data.qpos[:] = qpos0
        if qvel0 is not None:
            physics.data.qvel[:] = qvel0
        physics.forward()
    iteration, sim_iteration = 0, 0
    for _ in range(1, n_iterations*substeps):
        assert iteration < n_iterations
        full_step = not sim_iteration % substeps
        index = iteration % n_iterations
        if full_step or swim:
            physics2data(physics, index, data, maps, units, links_only=not full_step)
        if swim:
            handler.step(index)
            apply_xfrc(physics, data, index, maps['sensors'], units)
        if full_step and controller is not None:
            ctrl = controller(iteration, iteration*timestep)
            if ctrl is not None:
                physics.data.ctrl[:] = ctrl
        for _ in range(n_sub_steps):
            physics.step()
        sim_iteration += 1
        if not (sim_iteration + 1) % substeps:
            iteration += 1
    if iteration < n_iterations:
        physics2data(physics, iteration, data, maps, units)
        if swim:
            handler.step(iteration)
    return data, (physics.data.qpos.copy(), physics.data.qvel.copy())


# --------------------------------------------------------------------------
# the same loop without the interpreter (oracle/farms_loop.c) -- the CPU baseline
# --------------------------------------------------------------------------

class CompiledRollout:
    """``reference_rollout`` on oracle/farms_loop.c: physics2data, cycontacts2data, the swimming
    callback, a travelling-wave controller and ``mj_step`` in one C call per K iterations, as the
    reference's compiled glue (Cython -O3, NumPy kernels) would run them.  The NumPy functions
    above remain the checker (tests/test_oracle_farms_loop.py compares the two)."""
    # pylint: disable=too-many-instance-attributes

    def __init__(self, physics, spec, tables, buffer_size, wave=None, env_phase=0.0, swimming=True):
        import ctypes as ct
        from farms_mujoco_b200 import cabi
        from oracle import oracle as orc
        self.ct, self.lib = ct, orc.lib()
        self.physics, self.tables = physics, tables
        self.buffer_size = int(buffer_size)
        self.data = AnimatData.from_sensors_names(
            timestep=physics.model.timestep, buffer_size=self.buffer_size, links=spec.links_names,
            joints=spec.joints_names, contacts=spec.contacts_names, xfrc=spec.xfrc_names)
        self._farms = cabi.farms_to_c(tables)
        self._wave, self._keep = None, []
        if wave is not None:
            acts, amp, freq, lag = wave
            wc = cabi.FbWaveController()
            arrays = [np.ascontiguousarray(acts, dtype=np.int32)] + [
                np.ascontiguousarray(a, dtype=np.float64) for a in (amp, freq, lag, np.zeros(len(acts)))]
            self._keep = arrays
            wc.n = len(arrays[0])
            wc.actuator = arrays[0].ctypes.data_as(ct.POINTER(ct.c_int32))
            for name, arr in zip(('amplitude', 'frequency', 'phase_lag', 'offset'), arrays[1:]):
                setattr(wc, name, arr.ctypes.data_as(orc.c_double_p))
            self._wave = wc
        self.env_phase = float(env_phase)
        self.swimming = bool(swimming and len(tables.swim_links_index))
        self._norm = np.zeros(max(1, tables.n_contacts))
        self.stage_seconds = np.zeros(4)          # physics2data, swimming, control, mj_step
        self.iteration = 0

    def _arrays(self):
        orc_p = self.ct.POINTER(self.ct.c_double)
        s = self.data.sensors
        return [np.ascontiguousarray(a).ctypes.data_as(orc_p) if a.size else None
                for a in (s.links.array, s.joints.array, s.contacts.array, s.xfrc.array)] + [
                    self._norm.ctypes.data_as(orc_p)]

    def run(self, n_iterations, timed=False):
        """``n_iterations`` x (sensors, swimming, control, mj_step) from the current iteration."""
        ct = self.ct
        stage = self.stage_seconds.ctypes.data_as(ct.POINTER(ct.c_double)) if timed else None
        self.lib.orc_farms_run(
            self.physics._cmodel.byref(), ct.byref(self.physics._d), self._farms.byref(),
            ct.byref(self._wave) if self._wave is not None else None, self.env_phase,
            self.iteration, int(n_iterations), self.buffer_size, *self._arrays(), int(self.swimming), stage)
        self.iteration += int(n_iterations)

    def sensors(self):
        """The before_step of the current iteration alone (the last row of a rollout)."""
        ct = self.ct
        self.lib.orc_farms_sensors(
            self.physics._cmodel.byref(), ct.byref(self.physics._d), self._farms.byref(), self.iteration,
            self.buffer_size, *self._arrays(), int(self.swimming), None)
