"""ctypes wrapper of the fp64 CPU oracle (oracle/mjstep_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mjstep_oracle.c.  Only tests/,
``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl
reference`` legs may import this module.  PARITY UNPINNED (MuJoCo is absent from
the reference tree and from this image).

``OraclePhysics`` is shaped like the part of dm_control's ``Physics`` the
reference touches: ``physics.data.{qpos,qvel,ctrl,
xfrc_applied,xpos,xquat,xipos,sensordata,contact}``, ``physics.model.*``,
``physics.reset(keyframe_id)``, ``physics.step()``, so that the NumPy
restatements in farms_oracle.py read like farms_mujoco/simulation/physics.py.
"""

import ctypes as ct
import os
import subprocess
from types import SimpleNamespace

import numpy as np

from farms_mujoco_b200 import cabi
from farms_mujoco_b200 import mjcf_subset as ms

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = ct.POINTER(ct.c_int32)
c_double_p = ct.POINTER(ct.c_double)

_DOUBLE_FIELDS_A = ['qpos', 'qvel', 'ctrl', 'xfrc_applied', 'qpos_spring']
_DOUBLE_FIELDS_B = ['xpos', 'xquat', 'xmat', 'xipos', 'ximat', 'xanchor', 'xaxis', 'geom_xpos',
                    'geom_xmat', 'subtree_com', 'cinert', 'crb', 'cdof', 'cdof_dot', 'cvel',
                    'qM', 'qLD', 'qLDiagInv']


class OrcData(ct.Structure):
    _fields_ = (
        [(n, c_double_p) for n in _DOUBLE_FIELDS_A]
        + [('time', ct.c_double)]
        + [(n, c_double_p) for n in _DOUBLE_FIELDS_B]
        + [('ncon', ct.c_int32), ('con_cand', c_int_p), ('con_efc_address', c_int_p),
           ('con_dist', c_double_p), ('con_pos', c_double_p), ('con_frame', c_double_p),
           ('con_force', c_double_p),
           ('nefc', ct.c_int32), ('efc_type', c_int_p), ('efc_id', c_int_p),
           ('jnt_limit_row', c_int_p),
           ('efc_J', c_double_p), ('efc_pos', c_double_p), ('efc_margin', c_double_p),
           ('efc_R', c_double_p), ('efc_D', c_double_p), ('efc_KBIP', c_double_p),
           ('efc_aref', c_double_p), ('efc_force', c_double_p),
           ('qfrc_bias', c_double_p), ('qfrc_passive', c_double_p), ('qfrc_actuator', c_double_p),
           ('qfrc_smooth', c_double_p), ('qacc_smooth', c_double_p), ('qacc', c_double_p),
           ('qfrc_constraint', c_double_p), ('actuator_force', c_double_p),
           ('body_linvel', c_double_p), ('body_angvel', c_double_p),
           ('jnt_limit_force', c_double_p), ('solver_niter', ct.c_int32)]
    )


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, 'liboracle.so')
    srcs = [os.path.join(_HERE, name) for name in ('mjstep_oracle.c', 'farms_loop.c', 'oracle.h', 'Makefile')]
    if force or not os.path.exists(so) or any(
            os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src) for src in srcs):
        subprocess.run(['make', '-C', _HERE, '-B' if force else '-s'], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB  # pylint: disable=global-statement
    if _LIB is None:
        _LIB = ct.CDLL(build())
        for name in ('orc_kinematics', 'orc_com_pos', 'orc_crb', 'orc_collision',
                     'orc_make_constraint', 'orc_com_vel', 'orc_passive', 'orc_actuation',
                     'orc_solve_constraints', 'orc_forward', 'orc_euler', 'orc_step'):
            getattr(_LIB, name).argtypes = [ct.c_void_p, ct.c_void_p]
            getattr(_LIB, name).restype = None
        _LIB.orc_step_n.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int]
        _LIB.orc_step_n.restype = None
        # farms_loop.c: (model, data, farms, iteration, buffer_size, links, joints, contacts, xfrc,
        #                norm_scratch, swimming, stage_seconds[4] or NULL)
        _LIB.orc_farms_sensors.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_longlong, ct.c_int,
                                           c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                           ct.c_int, c_double_p]
        _LIB.orc_farms_sensors.restype = None
        # (model, data, farms, wave controller or NULL, env_phase, it0, n_iterations, buffer_size, ...)
        _LIB.orc_farms_run.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_double,
                                       ct.c_longlong, ct.c_int, ct.c_int,
                                       c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       ct.c_int, c_double_p]
        _LIB.orc_farms_run.restype = None
    return _LIB


class _Contact:
    """One ``physics.data.contact[i]`` entry."""

    def __init__(self, geom1, geom2, pos, frame, dist):
        self.geom1, self.geom2 = int(geom1), int(geom2)
        self.pos, self.frame, self.dist = pos, frame, dist


class OraclePhysics:
    """fp64 single-environment physics with a dm_control-``Physics``-like face."""
    # pylint: disable=too-many-instance-attributes

    def __init__(self, model):
        self.model = model
        self._cmodel = cabi.model_to_c(model)
        m = model
        nb, nj, nv, ng = m.nbody, m.njnt, m.nv, m.ngeom
        self.maxefc = max(1, 2*nj + 4*m.ncand)
        sizes = dict(
            qpos=m.nq, qvel=nv, ctrl=max(1, m.nu), xfrc_applied=6*nb, qpos_spring=m.nq,
            xpos=3*nb, xquat=4*nb, xmat=9*nb, xipos=3*nb, ximat=9*nb, xanchor=3*max(1, nj),
            xaxis=3*max(1, nj), geom_xpos=3*max(1, ng), geom_xmat=9*max(1, ng),
            subtree_com=3*nb, cinert=10*nb, crb=10*nb, cdof=6*nv, cdof_dot=6*nv, cvel=6*nb,
            qM=m.nM, qLD=m.nM, qLDiagInv=nv,
            con_dist=max(1, m.ncand), con_pos=3*max(1, m.ncand), con_frame=9*max(1, m.ncand),
            con_force=6*max(1, m.ncand),
            efc_J=self.maxefc*nv, efc_pos=self.maxefc, efc_margin=self.maxefc,
            efc_R=self.maxefc, efc_D=self.maxefc, efc_KBIP=4*self.maxefc,
            efc_aref=self.maxefc, efc_force=self.maxefc,
            qfrc_bias=nv, qfrc_passive=nv, qfrc_actuator=nv, qfrc_smooth=nv, qacc_smooth=nv,
            qacc=nv, qfrc_constraint=nv, actuator_force=max(1, m.nu),
            body_linvel=3*nb, body_angvel=3*nb, jnt_limit_force=max(1, nj),
        )
        isizes = dict(con_cand=max(1, m.ncand), con_efc_address=max(1, m.ncand),
                      efc_type=self.maxefc, efc_id=self.maxefc, jnt_limit_row=max(1, nj))
        self._d = OrcData()
        self.arrays = {}
        for name, size in sizes.items():
            arr = np.zeros(size, dtype=np.float64)
            self.arrays[name] = arr
            setattr(self._d, name, arr.ctypes.data_as(c_double_p))
        for name, size in isizes.items():
            arr = np.zeros(size, dtype=np.int32)
            self.arrays[name] = arr
            setattr(self._d, name, arr.ctypes.data_as(c_int_p))
        a = self.arrays
        self.data = SimpleNamespace(
            qpos=a['qpos'], qvel=a['qvel'], ctrl=a['ctrl'][:m.nu],
            xfrc_applied=a['xfrc_applied'].reshape(nb, 6),
            xpos=a['xpos'].reshape(nb, 3), xquat=a['xquat'].reshape(nb, 4),
            xmat=a['xmat'].reshape(nb, 9), xipos=a['xipos'].reshape(nb, 3),
            cvel=a['cvel'].reshape(nb, 6), qacc=a['qacc'],
            actuator_force=a['actuator_force'][:m.nu],
            sensordata=np.zeros(m.nsensordata), contact=[],
        )
        a['qpos_spring'][:] = m.qpos_spring
        self.reset(keyframe_id=None)

    # -- dm_control-like API -------------------------------------------------
    def reset(self, keyframe_id=0):
        """``physics.reset(keyframe_id)`` + the ``mj_forward`` of reset_context."""
        m, a = self.model, self.arrays
        a['qpos'][:] = m.qpos0 if keyframe_id is None else m.key_qpos
        a['qvel'][:] = 0 if keyframe_id is None else m.key_qvel
        a['ctrl'][:] = 0
        if keyframe_id is not None and m.nu:
            a['ctrl'][:m.nu] = m.key_ctrl
        a['xfrc_applied'][:] = 0
        self._d.time = 0.0
        self.forward()

    def forward(self):
        lib().orc_forward(self._cmodel.byref(), ct.byref(self._d))
        self._refresh()

    def step(self, n_sub_steps=1):
        """``mj_step`` x n: forward at the current state, then Euler."""
        for _ in range(n_sub_steps):
            lib().orc_step(self._cmodel.byref(), ct.byref(self._d))
        self._refresh()

    def step_raw(self, n):
        """n steps without refreshing the Python-side views (CPU-baseline timing)."""
        lib().orc_step_n(self._cmodel.byref(), ct.byref(self._d), n)

    # -- derived views --------------------------------------------------------
    @property
    def ncon(self):
        return int(self._d.ncon)

    @property
    def nefc(self):
        return int(self._d.nefc)

    @property
    def solver_niter(self):
        return int(self._d.solver_niter)

    @property
    def time(self):
        return float(self._d.time)

    def contact_force(self, i):
        """``mj_contactForce``: contact-frame (normal, t1, t2, 0, 0, 0)."""
        return self.arrays['con_force'].reshape(-1, 6)[i].copy()

    def full_mass_matrix(self):
        nv = self.model.nv
        out = np.zeros((nv, nv))
        lib().orc_fullM.argtypes = [ct.c_void_p, c_double_p, c_double_p]
        lib().orc_fullM(self._cmodel.byref(), self._d.qM, out.ctypes.data_as(c_double_p))
        return out

    def efc(self):
        n, nv = self.nefc, self.model.nv
        a = self.arrays
        return dict(J=a['efc_J'][:n*nv].reshape(n, nv).copy(), pos=a['efc_pos'][:n].copy(),
                    R=a['efc_R'][:n].copy(), D=a['efc_D'][:n].copy(),
                    aref=a['efc_aref'][:n].copy(), force=a['efc_force'][:n].copy(),
                    type=a['efc_type'][:n].copy(), id=a['efc_id'][:n].copy())

    def _refresh(self):
        m, a, d = self.model, self.arrays, self.data
        # sensordata in sensor order
        linvel = a['body_linvel'].reshape(-1, 3)
        angvel = a['body_angvel'].reshape(-1, 3)
        for sid in range(len(m.sensor_names)):
            stype, obj = m.sensor_type[sid], m.sensor_objid[sid]
            adr, dim = m.sensor_adr[sid], m.sensor_dim[sid]
            if stype == ms.SENS_FRAMEPOS:
                val = d.xipos[obj]
            elif stype == ms.SENS_FRAMEQUAT:
                val = ms.quat_mul(d.xquat[obj], m.body_iquat[obj])
            elif stype == ms.SENS_FRAMELINVEL:
                val = linvel[obj]
            elif stype == ms.SENS_FRAMEANGVEL:
                val = angvel[obj]
            elif stype == ms.SENS_JOINTPOS:
                val = a['qpos'][m.jnt_qposadr[obj]]
            elif stype == ms.SENS_JOINTVEL:
                val = a['qvel'][m.jnt_dofadr[obj]]
            elif stype == ms.SENS_JOINTLIMITFRC:
                val = a['jnt_limit_force'][obj]
            else:
                val = a['actuator_force'][obj]
            d.sensordata[adr:adr+dim] = val
        pos = a['con_pos'].reshape(-1, 3)
        frame = a['con_frame'].reshape(-1, 9)
        d.contact = [
            _Contact(m.cand_geom1[c], m.cand_geom2[c], pos[i].copy(), frame[i].copy(),
                     a['con_dist'][i])
            for i, c in enumerate(a['con_cand'][:self.ncon])
        ]
