"""ctypes mirror of include/farms_b200.h (struct definitions + marshalling).

The reference-side binding a maintainer would add is the same ctypes code
(INTEGRATION.md); the reference itself reaches native code through Cython
(farms_mujoco/sensors/sensors.pxd:12-29) and pybind (mujoco).
"""

import ctypes as ct

import numpy as np

from .layout import sc

c_int_p = ct.POINTER(ct.c_int32)
c_double_p = ct.POINTER(ct.c_double)
c_float_p = ct.POINTER(ct.c_float)
c_int64_p = ct.POINTER(ct.c_int64)


class FbModel(ct.Structure):
    _fields_ = [
        ('nbody', ct.c_int32), ('njnt', ct.c_int32), ('nq', ct.c_int32), ('nv', ct.c_int32),
        ('nu', ct.c_int32), ('ngeom', ct.c_int32), ('ncand', ct.c_int32), ('nM', ct.c_int32),
        ('timestep', ct.c_double), ('gravity', ct.c_double*3), ('impratio', ct.c_double),
        ('solver_iterations', ct.c_int32), ('tolerance', ct.c_double), ('meaninertia', ct.c_double),
        ('body_parentid', c_int_p), ('body_jntid', c_int_p), ('body_dofadr', c_int_p),
        ('body_dofnum', c_int_p),
        ('body_pos', c_double_p), ('body_quat', c_double_p), ('body_ipos', c_double_p),
        ('body_iquat', c_double_p),
        ('body_mass', c_double_p), ('body_inertia', c_double_p), ('body_invweight0', c_double_p),
        ('jnt_type', c_int_p), ('jnt_bodyid', c_int_p), ('jnt_qposadr', c_int_p),
        ('jnt_dofadr', c_int_p), ('jnt_limited', c_int_p),
        ('jnt_pos', c_double_p), ('jnt_axis', c_double_p), ('jnt_stiffness', c_double_p),
        ('jnt_range', c_double_p), ('jnt_margin', c_double_p),
        ('jnt_solref', c_double_p), ('jnt_solimp', c_double_p),
        ('dof_bodyid', c_int_p), ('dof_jntid', c_int_p), ('dof_parentid', c_int_p),
        ('dof_Madr', c_int_p),
        ('dof_damping', c_double_p), ('dof_armature', c_double_p), ('dof_invweight0', c_double_p),
        ('qpos0', c_double_p), ('qpos_spring', c_double_p),
        ('geom_type', c_int_p), ('geom_bodyid', c_int_p),
        ('geom_pos', c_double_p), ('geom_quat', c_double_p), ('geom_size', c_double_p),
        ('cand_geom1', c_int_p), ('cand_geom2', c_int_p), ('cand_end', c_int_p),
        ('cand_friction', c_double_p), ('cand_solref', c_double_p), ('cand_solimp', c_double_p),
        ('cand_margin', c_double_p), ('cand_gap', c_double_p),
        ('actuator_trnid', c_int_p), ('actuator_ctrllimited', c_int_p),
        ('actuator_forcelimited', c_int_p),
        ('actuator_gainprm', c_double_p), ('actuator_biasprm', c_double_p),
        ('actuator_ctrlrange', c_double_p), ('actuator_forcerange', c_double_p),
        ('actuator_gear', c_double_p),
        ('key_qpos', c_double_p), ('key_qvel', c_double_p),
    ]


class FbFarms(ct.Structure):
    _fields_ = [
        ('n_links', ct.c_int32), ('n_joints', ct.c_int32), ('n_contacts', ct.c_int32),
        ('n_xfrc', ct.c_int32), ('n_swim', ct.c_int32),
        ('link_cols', ct.c_int32), ('joint_cols', ct.c_int32), ('contact_cols', ct.c_int32),
        ('xfrc_cols', ct.c_int32),
        ('col_joint_position', ct.c_int32), ('col_joint_velocity', ct.c_int32),
        ('col_joint_torque', ct.c_int32), ('col_joint_limit_force', ct.c_int32),
        ('link_body', c_int_p), ('joint_qposadr', c_int_p), ('joint_dofadr', c_int_p),
        ('joint_jntid', c_int_p), ('joint_act_position', c_int_p),
        ('joint_act_velocity', c_int_p), ('joint_act_torque', c_int_p),
        ('cand_sensor', c_int_p), ('xfrc_body', c_int_p),
        ('swim_links_index', c_int_p), ('swim_xfrc_index', c_int_p),
        ('swim_mass', c_double_p), ('swim_height', c_double_p), ('swim_density', c_double_p),
        ('swim_coefficients', c_double_p),
        ('water_drag', ct.c_int32), ('water_sph', ct.c_int32), ('water_buoyancy', ct.c_int32),
        ('water_surface', ct.c_double), ('water_density', ct.c_double),
        ('water_viscosity', ct.c_double), ('water_velocity', ct.c_double*3),
        ('meters', ct.c_double), ('seconds', ct.c_double), ('kilograms', ct.c_double),
    ]


class FbWaveController(ct.Structure):
    _fields_ = [
        ('n', ct.c_int32), ('actuator', c_int_p),
        ('amplitude', c_double_p), ('frequency', c_double_p), ('phase_lag', c_double_p),
        ('offset', c_double_p),
    ]


class FbCpgNetwork(ct.Structure):
    _fields_ = [
        ('n_osc', ct.c_int32), ('n_coupling', ct.c_int32), ('n_out', ct.c_int32),
        ('frequency', c_double_p), ('amplitude', c_double_p), ('rate', c_double_p),
        ('coupling_from', c_int_p), ('coupling_to', c_int_p),
        ('coupling_weight', c_double_p), ('coupling_bias', c_double_p),
        ('out_actuator', c_int_p), ('out_osc_a', c_int_p), ('out_osc_b', c_int_p),
        ('out_gain', c_double_p), ('out_offset', c_double_p),
    ]


class FbLogView(ct.Structure):
    _fields_ = [
        ('links_dev', c_float_p), ('joints_dev', c_float_p), ('contacts_dev', c_float_p),
        ('xfrc_dev', c_float_p),
        ('links_vec', ct.c_int32), ('joints_vec', ct.c_int32),
        ('contacts_vec', ct.c_int32), ('xfrc_vec', ct.c_int32),
        ('ring', ct.c_int32), ('n_envs', ct.c_int32), ('env_pad', ct.c_int32),
    ]


class FbStateView(ct.Structure):
    _fields_ = [
        ('qpos_dev', c_float_p), ('qvel_dev', c_float_p), ('ctrl_dev', c_float_p),
        ('xfrc_applied_dev', c_float_p), ('qpos_spring_dev', c_float_p),
        ('env_phase_dev', c_float_p), ('flags_dev', c_int_p), ('iteration_dev', c_int64_p),
    ]


class FbDerivedView(ct.Structure):
    _fields_ = [
        ('xpos_dev', c_float_p), ('xquat_dev', c_float_p), ('xipos_dev', c_float_p),
        ('linvel_dev', c_float_p), ('angvel_dev', c_float_p),
        ('actuator_force_dev', c_float_p), ('jnt_limit_force_dev', c_float_p),
        ('qacc_dev', c_float_p),
        ('ncon_dev', c_int_p), ('con_cand_dev', c_int_p),
        ('con_dist_dev', c_float_p), ('con_pos_dev', c_float_p), ('con_frame_dev', c_float_p),
        ('con_force_dev', c_float_p), ('maxcon', ct.c_int32),
    ]


def _arr(values, dtype):
    return np.ascontiguousarray(np.asarray(values, dtype=dtype))


class Marshalled:
    """A ctypes struct plus the NumPy arrays that keep its pointers alive."""

    def __init__(self, struct):
        self.struct = struct
        self.keep = {}

    def set_int(self, name, values):
        arr = _arr(values, np.int32).ravel()
        if arr.size == 0:
            arr = np.zeros(1, dtype=np.int32)
        self.keep[name] = arr
        setattr(self.struct, name, arr.ctypes.data_as(c_int_p))

    def set_double(self, name, values):
        arr = _arr(values, np.float64).ravel()
        if arr.size == 0:
            arr = np.zeros(1, dtype=np.float64)
        self.keep[name] = arr
        setattr(self.struct, name, arr.ctypes.data_as(c_double_p))

    def byref(self):
        return ct.byref(self.struct)


def model_to_c(model, qpos_spring=None):
    """Flat ``Model`` (mjcf_subset.py) -> ``FbModel``."""
    out = Marshalled(FbModel())
    s = out.struct
    s.nbody, s.njnt, s.nq, s.nv = model.nbody, model.njnt, model.nq, model.nv
    s.nu, s.ngeom, s.ncand, s.nM = model.nu, model.ngeom, model.ncand, model.nM
    s.timestep = model.timestep
    s.gravity = (ct.c_double*3)(*[float(g) for g in model.gravity])
    s.impratio = model.impratio
    s.solver_iterations = model.iterations
    s.tolerance = model.tolerance
    s.meaninertia = model.meaninertia
    for name in ('body_parentid', 'body_jntid', 'body_dofadr', 'body_dofnum',
                 'jnt_type', 'jnt_bodyid', 'jnt_qposadr', 'jnt_dofadr', 'jnt_limited',
                 'dof_bodyid', 'dof_jntid', 'dof_parentid', 'dof_Madr',
                 'geom_type', 'geom_bodyid', 'cand_geom1', 'cand_geom2', 'cand_end',
                 'actuator_trnid', 'actuator_ctrllimited', 'actuator_forcelimited'):
        out.set_int(name, getattr(model, name))
    for name in ('body_pos', 'body_quat', 'body_ipos', 'body_iquat', 'body_mass',
                 'body_inertia', 'body_invweight0',
                 'jnt_pos', 'jnt_axis', 'jnt_stiffness', 'jnt_range', 'jnt_margin',
                 'jnt_solref', 'jnt_solimp',
                 'dof_damping', 'dof_armature', 'dof_invweight0', 'qpos0',
                 'geom_pos', 'geom_quat', 'geom_size',
                 'cand_friction', 'cand_solref', 'cand_solimp', 'cand_margin', 'cand_gap',
                 'actuator_gainprm', 'actuator_biasprm', 'actuator_ctrlrange',
                 'actuator_forcerange', 'actuator_gear', 'key_qpos', 'key_qvel'):
        out.set_double(name, getattr(model, name))
    out.set_double('qpos_spring', model.qpos_spring if qpos_spring is None else qpos_spring)
    return out


def farms_to_c(tables):
    """``FarmsTables`` (simulation/physics.py) -> ``FbFarms``."""
    out = Marshalled(FbFarms())
    s = out.struct
    s.n_links, s.n_joints = len(tables.link_body), len(tables.joint_qposadr)
    s.n_contacts, s.n_xfrc = tables.n_contacts, len(tables.xfrc_body)
    s.n_swim = len(tables.swim_links_index)
    s.link_cols, s.joint_cols = sc.link_size, sc.joint_size
    s.contact_cols, s.xfrc_cols = sc.contact_size, sc.xfrc_size
    s.col_joint_position = sc.joint_position
    s.col_joint_velocity = sc.joint_velocity
    s.col_joint_torque = sc.joint_torque
    s.col_joint_limit_force = sc.joint_limit_force
    for name in ('link_body', 'joint_qposadr', 'joint_dofadr', 'joint_jntid',
                 'joint_act_position', 'joint_act_velocity', 'joint_act_torque',
                 'cand_sensor', 'xfrc_body', 'swim_links_index', 'swim_xfrc_index'):
        out.set_int(name, getattr(tables, name))
    for name in ('swim_mass', 'swim_height', 'swim_density', 'swim_coefficients'):
        out.set_double(name, getattr(tables, name))
    s.water_drag, s.water_sph = int(tables.water_drag), int(tables.water_sph)
    s.water_buoyancy = int(tables.water_buoyancy)
    s.water_surface = tables.water_surface
    s.water_density = tables.water_density
    s.water_viscosity = tables.water_viscosity
    s.water_velocity = (ct.c_double*3)(*[float(v) for v in tables.water_velocity])
    s.meters, s.seconds, s.kilograms = tables.meters, tables.seconds, tables.kilograms
    return out
