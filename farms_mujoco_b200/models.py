"""Synthetic FARMS-schema animats (SWIMMER8, SALAMANDER, CENTIPEDE).

The reference builds its MJCF from SDF files through dm_control
(farms_mujoco/simulation/mjcf.py:647-1035, 1174-1512); neither the SDF assets
nor dm_control exist here, so these generators emit MJCF *text* that follows
the same schema and naming rules: a model-root body
carrying ``<freejoint name="root_<model>">``, nested link bodies with one
hinge each, ``contype=1 conaffinity=0 condim=3 margin=0 group=2`` collision
geoms, ``<inertial pos mass fullinertia>``, a position/velocity/motor actuator
triple per controlled joint, framelinvel/frameangvel + jointpos/jointvel/
jointlimitfrc + actuatorfrc sensors, one keyframe "initial", and the
compiler/size/option blocks of mjcf.py:1245-1403.  They are the workloads of
BASELINE.json ``configs``.
"""

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .mjcf_subset import euler_xyz2quat, quat_mul, axisangle2quat
from .options import (
    AnimatOptions, ArenaOptions, SimulationOptions, LinkOptions, JointOptions,
    MotorOptions, MorphologyOptions, ControlOptions, SpawnOptions, WaterOptions,
)


def _fmt(values):
    return ' '.join(repr(float(v)) for v in values)


def capsule_mass_inertia(radius, length, density=1000.0):
    """Mass and (axial, transverse) inertia of a solid capsule about its centre."""
    m_cyl = density*np.pi*radius**2*length
    m_sph = density*4.0/3.0*np.pi*radius**3
    axial = 0.5*m_cyl*radius**2 + 0.4*m_sph*radius**2
    transverse = (
        m_cyl*(length**2/12.0 + radius**2/4.0)
        + m_sph*(0.4*radius**2 + length**2/4.0 + 3.0*length*radius/8.0)
    )
    return m_cyl + m_sph, axial, transverse


def sphere_mass_inertia(radius, density=1000.0):
    mass = density*4.0/3.0*np.pi*radius**3
    return mass, 0.4*mass*radius**2


def _quat_z_to(direction):
    """Quaternion (wxyz) rotating the local z axis onto ``direction``."""
    d = np.asarray(direction, dtype=float)
    d = d/np.linalg.norm(d)
    z = np.array([0.0, 0.0, 1.0])
    cross = np.cross(z, d)
    s = np.linalg.norm(cross)
    if s < 1e-12:
        return np.array([1.0, 0, 0, 0]) if d[2] > 0 else np.array([0.0, 1, 0, 0])
    return axisangle2quat(cross/s, np.arctan2(s, z @ d))


@dataclass
class _Geom:
    name: str
    type: str
    size: Tuple[float, float, float]
    pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    quat: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 0.0)


@dataclass
class _Link:
    name: str
    parent: str            # '' -> child of the model-root body
    pos: Tuple[float, float, float]
    joint: str = ''        # '' -> welded (the base link)
    axis: Tuple[float, float, float] = (0.0, 0.0, 1.0)
    limits: Tuple[float, float] = (-1.0, 1.0)
    mass: float = 0.0
    ipos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    inertia: Tuple[float, float, float] = (0.0, 0.0, 0.0)   # diagonal, body axes
    geoms: List[_Geom] = field(default_factory=list)
    offdiag: Tuple[float, float, float] = (0.0, 0.0, 0.0)   # Ixy, Ixz, Iyz
    jpos: Tuple[float, float, float] = (0.0, 0.0, 0.0)      # joint anchor in the link frame (SDF joint pose)
    quat: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 0.0)   # link frame in the parent's (wxyz)
    jtype: str = 'hinge'                                     # 'hinge' or 'slide' (SDF prismatic)


@dataclass
class AnimatSpec:
    """Everything a simulation needs: MJCF text + the option objects."""
    name: str
    mjcf: str
    animat_options: AnimatOptions
    arena_options: ArenaOptions
    simulation_options: SimulationOptions
    links_names: List[str]
    joints_names: List[str]
    contacts_names: List[Tuple[str, str]]
    xfrc_names: List[str]
    base_link: str


def _emit_mjcf(model_name, links, joints_opts, motors, spawn_pose, sim, arena_z,
               water_height, friction, self_collisions=()):
    # pylint: disable=too-many-locals,too-many-arguments
    children = {}
    for link in links:
        children.setdefault(link.parent, []).append(link)
    joint_order = []
    out = []

    def emit_link(link, indent):
        pad = ' '*indent
        out.append(f'{pad}<body name="{link.name}" pos="{_fmt(link.pos)}" quat="{_fmt(link.quat)}">')
        if link.joint:
            jo = joints_opts[link.joint]
            joint_order.append(link.joint)
            out.append(
                f'{pad}  <joint name="{link.joint}" type="{link.jtype}" axis="{_fmt(link.axis)}" '
                f'pos="{_fmt(link.jpos)}" damping="{float(jo.damping)!r}" '
                f'stiffness="{float(jo.stiffness)!r}" springref="0.0" frictionloss="0.0" '
                f'limited="true" range="{_fmt(link.limits)}"/>')
        for geom in link.geoms:
            out.append(
                f'{pad}  <geom name="{geom.name}" type="{geom.type}" size="{_fmt(geom.size)}" '
                f'pos="{_fmt(geom.pos)}" quat="{_fmt(geom.quat)}" friction="{_fmt(friction)}" '
                f'margin="0.0" contype="1" conaffinity="0" condim="3" group="2"/>')
            # visual twin (mjcf.py:240-250); discarded by the front-end
            out.append(
                f'{pad}  <geom name="{geom.name}_visual" type="{geom.type}" '
                f'size="{_fmt(geom.size)}" pos="{_fmt(geom.pos)}" quat="{_fmt(geom.quat)}" '
                f'contype="0" conaffinity="0" group="1"/>')
        ixx, iyy, izz = link.inertia
        out.append(
            f'{pad}  <inertial pos="{_fmt(link.ipos)}" mass="{float(link.mass)!r}" '
            f'fullinertia="{_fmt([ixx, iyy, izz, *link.offdiag])}"/>')
        for child in children.get(link.name, []):
            emit_link(child, indent + 2)
        out.append(f'{pad}</body>')

    spawn_quat = euler_xyz2quat(spawn_pose[3:])
    out.append(f'<mujoco model="{model_name}">')
    out.append(
        '  <compiler angle="radian" eulerseq="xyz" boundmass="0" boundinertia="0" '
        'balanceinertia="false" inertiafromgeom="false" fusestatic="true" '
        'discardvisual="true"/>')
    out.append('  <size nkey="1" njmax="4096" nconmax="4096"/>')
    out.append(
        f'  <option timestep="{sim.timestep/max(1, sim.num_sub_steps)!r}" impratio="{sim.impratio!r}" '
        f'gravity="{_fmt(sim.gravity)}" cone="{sim.cone}" solver="{sim.solver}" '
        f'iterations="{sim.n_solver_iters}" integrator="{sim.integrator}"/>')
    out.append('  <worldbody>')
    # arena (fixed base, friction 0, all_collisions -> conaffinity=1; mjcf.py:1195-1212)
    out.append(f'    <body name="arena" pos="0.0 0.0 {float(arena_z)!r}" quat="1.0 0.0 0.0 0.0">')
    out.append(
        '      <geom name="floor" type="plane" size="10.0 10.0 0.1" pos="0.0 0.0 0.0" '
        'quat="1.0 0.0 0.0 0.0" friction="0.0 0.0 0.0" margin="0.0" contype="1" '
        'conaffinity="1" condim="3" group="2"/>')
    out.append('    </body>')
    if water_height is not None:
        out.append(f'    <body name="water" pos="0.0 0.0 {float(water_height)!r}" quat="1.0 0.0 0.0 0.0">')
        out.append(
            '      <geom name="water_surface" type="plane" size="10.0 10.0 0.1" '
            'contype="0" conaffinity="0" group="1"/>')
        out.append('    </body>')
    out.append(
        f'    <body name="{model_name}" pos="{_fmt(spawn_pose[:3])}" quat="{_fmt(spawn_quat)}">')
    out.append(f'      <freejoint name="root_{model_name}"/>')
    for link in children.get('', []):
        emit_link(link, 6)
    out.append('    </body>')
    out.append('  </worldbody>')
    # actuators (mjcf.py:819-866)
    out.append('  <actuator>')
    for motor in motors:
        j = motor.joint_name
        out.append(
            f'    <position name="actuator_position_{j}" joint="{j}" kp="{float(motor.gains[0])!r}" '
            'ctrllimited="false" ctrlrange="-1000000.0 1000000.0" forcelimited="false" '
            'forcerange="-1000000.0 1000000.0"/>')
        out.append(
            f'    <velocity name="actuator_velocity_{j}" joint="{j}" kv="{float(motor.gains[1])!r}" '
            'ctrllimited="false" ctrlrange="-1000000.0 1000000.0" forcelimited="false" '
            'forcerange="-1000000.0 1000000.0"/>')
        out.append(f'    <motor name="actuator_torque_{j}" joint="{j}"/>')
    out.append('  </actuator>')
    # sensors (mjcf.py:950-1009): links incl. the model-root body, then joints, then actuators
    out.append('  <sensor>')
    for name in [model_name] + [link.name for link in links]:
        out.append(f'    <framelinvel name="framelinvel_{name}" objname="{name}" objtype="body"/>')
        out.append(f'    <frameangvel name="frameangvel_{name}" objname="{name}" objtype="body"/>')
    for link in links:
        if link.joint:
            j = link.joint
            out.append(f'    <jointpos name="jointpos_{j}" joint="{j}"/>')
            out.append(f'    <jointvel name="jointvel_{j}" joint="{j}"/>')
            out.append(f'    <jointlimitfrc name="jointlimitfrc_{j}" joint="{j}"/>')
    for motor in motors:
        j = motor.joint_name
        for tag, kind in (('position', 'position'), ('velocity', 'velocity'), ('motor', 'torque')):
            out.append(
                f'    <actuatorfrc name="actuatorfrc_{tag}_{j}" actuator="actuator_{kind}_{j}"/>')
    out.append('  </sensor>')
    # explicit self-collision pairs (mjcf.py:1012-1033): one per couple of collision geoms of the
    # two links, condim 3, friction [0]*5
    if self_collisions:
        collision_map = {link.name: [geom.name for geom in link.geoms] for link in links}
        out.append('  <contact>')
        for pair_i, (link1, link2) in enumerate(self_collisions):
            for col1_i, col1_name in enumerate(collision_map[link1]):
                for col2_i, col2_name in enumerate(collision_map[link2]):
                    out.append(f'    <pair name="contact_pair_{pair_i}_{col1_i}_{col2_i}" geom1="{col1_name}" '
                               f'geom2="{col2_name}" condim="3" friction="0 0 0 0 0"/>')
        out.append('  </contact>')
    # keyframe (mjcf.py:743-788)
    qpos = list(spawn_pose[:3]) + list(spawn_quat) + [joints_opts[j].initial[0] for j in joint_order]
    qvel = [0.0]*6 + [joints_opts[j].initial[1] for j in joint_order]
    out.append('  <keyframe>')
    out.append(f'    <key name="initial" time="0.0" qpos="{_fmt(qpos)}" qvel="{_fmt(qvel)}"/>')
    out.append('  </keyframe>')
    out.append('</mujoco>')
    return '\n'.join(out) + '\n', joint_order


def _x_capsule(name, radius, length):
    """Capsule along the link's x axis, proximal end at the link origin."""
    return _Geom(name=name, type='capsule', size=(radius, 0.5*length, 0.0),
                 pos=(0.5*length, 0.0, 0.0),
                 quat=tuple(euler_xyz2quat([0.0, 0.5*np.pi, 0.0])))


def _finish(name, links, joints_cfg, swimming, drag_coefficients, water, arena_z, spawn_pose,
            contacts_names, timestep, n_iterations, friction):
    # pylint: disable=too-many-arguments,too-many-locals
    sim = SimulationOptions(timestep=timestep, n_iterations=n_iterations)
    joints_opts = {
        jname: JointOptions(name=jname, initial=[0.0, 0.0], stiffness=cfg['stiffness'],
                            damping=cfg['damping'], limits=list(cfg['limits']))
        for jname, cfg in joints_cfg.items()
    }
    motors = [
        MotorOptions(joint_name=jname, control_types=['position'],
                     gains=[cfg['kp'], cfg['kv']])
        for jname, cfg in joints_cfg.items()
    ]
    mjcf, joint_order = _emit_mjcf(
        model_name=name, links=links, joints_opts=joints_opts, motors=motors,
        spawn_pose=spawn_pose, sim=sim, arena_z=arena_z,
        water_height=water.height, friction=friction)
    def coefficients(link):
        # a callable gives per-link coefficients (real FARMS animats scale them with the
        # link's size; explicit quadratic drag is unstable when c*|v|*dt/m approaches 1)
        coefs = drag_coefficients(link) if callable(drag_coefficients) else drag_coefficients
        return [list(coefs[0]), list(coefs[1])]

    links_opts = [
        LinkOptions(name=link.name, swimming=swimming, density=1000.0,
                    drag_coefficients=coefficients(link), friction=list(friction))
        for link in links
    ]
    animat_options = AnimatOptions(
        name=name,
        spawn=SpawnOptions(pose=list(spawn_pose)),
        morphology=MorphologyOptions(
            links=links_opts, joints=[joints_opts[j] for j in joint_order]),
        control=ControlOptions(motors=motors),
    )
    arena_options = ArenaOptions(ground_height=arena_z, water=water)
    return AnimatSpec(
        name=name, mjcf=mjcf, animat_options=animat_options, arena_options=arena_options,
        simulation_options=sim,
        links_names=[link.name for link in links],
        joints_names=joint_order,
        contacts_names=contacts_names,
        xfrc_names=[link.name for link in links] if swimming else [],
        base_link=links[0].name,
    )


def swimmer8(timestep=1e-3, n_iterations=10000):
    """BASELINE config 1: 8-link anguilliform swimmer, drag on."""
    radius, length = 0.02, 0.1
    mass, axial, transverse = capsule_mass_inertia(radius, length)
    links, joints_cfg = [], {}
    for i in range(8):
        joint = f'joint_{i-1}' if i else ''
        links.append(_Link(
            name=f'link_{i}', parent=f'link_{i-1}' if i else '',
            pos=(length, 0.0, 0.0) if i else (0.0, 0.0, 0.0),
            joint=joint, axis=(0.0, 0.0, 1.0), limits=(-1.0, 1.0),
            mass=mass, ipos=(0.5*length, 0.0, 0.0), inertia=(axial, transverse, transverse),
            geoms=[_x_capsule(f'link_{i}_collision', radius, length)],
        ))
        if joint:
            joints_cfg[joint] = dict(stiffness=0.0, damping=1e-3, limits=(-1.0, 1.0),
                                     kp=1.0, kv=1e-3)
    water = WaterOptions(height=0.0, drag=True, buoyancy=True, density=1000.0,
                         velocity=[0.0, 0.0, 0.0], viscosity=1.0)
    return _finish(
        'swimmer', links, joints_cfg, swimming=True,
        drag_coefficients=[[-0.1, -1.0, -1.0], [-1e-3, -1e-3, -1e-3]],
        water=water, arena_z=-2.0, spawn_pose=[0.0, 0.0, -0.1, 0.0, 0.0, 0.0],
        contacts_names=[(f'link_{i}', '') for i in range(8)],
        timestep=timestep, n_iterations=n_iterations, friction=[1.0, 0.0, 0.0])


def salamander(swimming=False, timestep=1e-3, n_iterations=1000):
    """BASELINE configs 2/3/5: 12 body links + 4 legs x 4 hinges (28 links, nv=33)."""
    # pylint: disable=too-many-locals
    n_body, seg_len = 12, 0.08
    links, joints_cfg = [], {}
    for i in range(n_body):
        radius = 0.025*(1.0 - 0.04*i)
        mass, axial, transverse = capsule_mass_inertia(radius, seg_len)
        joint = f'joint_body_{i-1}' if i else ''
        links.append(_Link(
            name=f'link_body_{i}', parent=f'link_body_{i-1}' if i else '',
            pos=(seg_len, 0.0, 0.0) if i else (0.0, 0.0, 0.0),
            joint=joint, axis=(0.0, 0.0, 1.0), limits=(-1.0, 1.0),
            mass=mass, ipos=(0.5*seg_len, 0.0, 0.0), inertia=(axial, transverse, transverse),
            geoms=[_x_capsule(f'link_body_{i}_collision', radius, seg_len)],
        ))
        if joint:
            joints_cfg[joint] = dict(stiffness=0.0, damping=5e-3, limits=(-1.0, 1.0),
                                     kp=2.0, kv=2e-2)
    r_leg, r_foot, limb = 0.008, 0.01, 0.04
    m_sph, i_sph = sphere_mass_inertia(r_leg)
    m_cap, ax_cap, tr_cap = capsule_mass_inertia(r_leg, limb)
    m_foot, i_foot = sphere_mass_inertia(r_foot)
    for leg_i, attach in enumerate((1, 5)):
        for side_i, (side, sgn) in enumerate((('L', 1.0), ('R', -1.0))):
            base = f'leg_{leg_i}_{side}'
            names = [f'link_{base}_{k}' for k in range(4)]
            jnames = [f'joint_{base}_{k}' for k in range(4)]
            # 0: yaw at the hip, 1: pitch (same origin), 2: roll + upper limb, 3: elbow + lower limb
            links.append(_Link(
                name=names[0], parent=f'link_body_{attach}', pos=(0.5*seg_len, sgn*0.03, 0.0),
                joint=jnames[0], axis=(0.0, 0.0, 1.0), limits=(-1.2, 1.2),
                mass=m_sph, inertia=(i_sph, i_sph, i_sph),
                geoms=[_Geom(f'{names[0]}_collision', 'sphere', (r_leg, r_leg, r_leg))]))
            links.append(_Link(
                name=names[1], parent=names[0], pos=(0.0, 0.0, 0.0),
                joint=jnames[1], axis=(1.0, 0.0, 0.0), limits=(-1.2, 1.2),
                mass=m_sph, inertia=(i_sph, i_sph, i_sph),
                geoms=[_Geom(f'{names[1]}_collision', 'sphere', (r_leg, r_leg, r_leg))]))
            links.append(_Link(
                name=names[2], parent=names[1], pos=(0.0, 0.0, 0.0),
                joint=jnames[2], axis=(0.0, 1.0, 0.0), limits=(-1.2, 1.2),
                mass=m_cap, ipos=(0.0, sgn*0.5*limb, 0.0), inertia=(tr_cap, ax_cap, tr_cap),
                geoms=[_Geom(f'{names[2]}_collision', 'capsule', (r_leg, 0.5*limb, 0.0),
                             pos=(0.0, sgn*0.5*limb, 0.0),
                             quat=tuple(_quat_z_to([0.0, sgn, 0.0])))]))
            # lower limb: capsule pointing down + foot sphere; composite inertia about the CoM
            m_low = m_cap + m_foot
            z_com = (m_cap*(-0.5*limb) + m_foot*(-limb))/m_low
            i_tr = (tr_cap + m_cap*(-0.5*limb - z_com)**2 + i_foot + m_foot*(-limb - z_com)**2)
            i_ax = ax_cap + i_foot
            links.append(_Link(
                name=names[3], parent=names[2], pos=(0.0, sgn*limb, 0.0),
                joint=jnames[3], axis=(1.0, 0.0, 0.0), limits=(-1.2, 1.2),
                mass=m_low, ipos=(0.0, 0.0, z_com), inertia=(i_tr, i_tr, i_ax),
                geoms=[
                    _Geom(f'{names[3]}_collision', 'capsule', (r_leg, 0.5*limb, 0.0),
                          pos=(0.0, 0.0, -0.5*limb)),
                    _Geom(f'{names[3]}_foot', 'sphere', (r_foot, r_foot, r_foot),
                          pos=(0.0, 0.0, -limb)),
                ]))
            for jname in jnames:
                joints_cfg[jname] = dict(stiffness=0.0, damping=1e-3, limits=(-1.2, 1.2),
                                         kp=0.5, kv=5e-3)
    contacts = [(f'link_body_{i}', '') for i in range(n_body)] + [
        (f'link_leg_{leg_i}_{side}_3', '') for leg_i in range(2) for side in ('L', 'R')]
    if swimming:
        water = WaterOptions(height=0.0, drag=True, buoyancy=True, viscosity=1.0)
        spawn = [0.0, 0.0, -0.2, 0.0, 0.0, 0.0]
        arena_z = -2.0
    else:
        water = WaterOptions(height=None)
        spawn = [0.0, 0.0, limb + r_foot + 0.002, 0.0, 0.0, 0.0]
        arena_z = 0.0
    def drag(link):
        if 'leg' in link.name:   # ~ frontal area / r^5 scaling of the 8 mm limb segments
            return [[-0.025, -0.25, -0.25], [-1e-8, -1e-8, -1e-8]]
        return [[-0.5, -5.0, -5.0], [-1e-3, -1e-3, -1e-3]]

    return _finish(
        'salamander', links, joints_cfg, swimming=swimming,
        drag_coefficients=drag,
        water=water, arena_z=arena_z, spawn_pose=spawn, contacts_names=contacts,
        timestep=timestep, n_iterations=n_iterations, friction=[1.0, 0.0, 0.0])


def centipede(timestep=1e-3, n_iterations=1000):
    """BASELINE config 4: 14 segments + 28 one-hinge capsule legs (42 links, nv=47)."""
    # pylint: disable=too-many-locals
    n_seg, seg_len, r_seg = 14, 0.05, 0.015
    leg_len, r_leg = 0.04, 0.005
    mass, axial, transverse = capsule_mass_inertia(r_seg, seg_len)
    m_leg, ax_leg, tr_leg = capsule_mass_inertia(r_leg, leg_len)
    links, joints_cfg = [], {}
    for i in range(n_seg):
        joint = f'joint_body_{i-1}' if i else ''
        links.append(_Link(
            name=f'link_body_{i}', parent=f'link_body_{i-1}' if i else '',
            pos=(seg_len, 0.0, 0.0) if i else (0.0, 0.0, 0.0),
            joint=joint, axis=(0.0, 0.0, 1.0), limits=(-0.8, 0.8),
            mass=mass, ipos=(0.5*seg_len, 0.0, 0.0), inertia=(axial, transverse, transverse),
            geoms=[_x_capsule(f'link_body_{i}_collision', r_seg, seg_len)]))
        if joint:
            joints_cfg[joint] = dict(stiffness=0.0, damping=2e-3, limits=(-0.8, 0.8),
                                     kp=1.0, kv=1e-2)
    for i in range(n_seg):
        for side, sgn in (('L', 1.0), ('R', -1.0)):
            direction = np.array([0.0, sgn*0.6, -0.8])
            quat = _quat_z_to(direction)
            # inertia of a capsule whose axis lies in the y-z plane, expressed in link axes
            rot = np.array([[1.0, 0, 0], [0, 0.8, sgn*0.6], [0, -sgn*0.6, 0.8]])  # cols: x, t, axis
            inertia_full = rot @ np.diag([tr_leg, tr_leg, ax_leg]) @ rot.T
            name = f'link_leg_{i}_{side}'
            jname = f'joint_leg_{i}_{side}'
            link = _Link(
                name=name, parent=f'link_body_{i}', pos=(0.5*seg_len, sgn*r_seg, 0.0),
                joint=jname, axis=(1.0, 0.0, 0.0), limits=(-0.6, 0.6),
                mass=m_leg, ipos=tuple(0.5*leg_len*direction),
                inertia=(inertia_full[0, 0], inertia_full[1, 1], inertia_full[2, 2]),
                geoms=[_Geom(f'{name}_collision', 'capsule', (r_leg, 0.5*leg_len, 0.0),
                             pos=tuple(0.5*leg_len*direction), quat=tuple(quat))],
                offdiag=(0.0, 0.0, inertia_full[1, 2]))
            links.append(link)
            # kv: MuJoCo's Euler integrator treats actuator velocity feedback explicitly, stable
            # while kv*dt/I < 2; the leg's inertia about its hinge is 2e-6 kg m^2, so kv = 3e-3
            # (1.5, plus the position actuator) diverged in ~1.5 % of the random rollouts, in
            # the fp64 oracle as well
            joints_cfg[jname] = dict(stiffness=0.0, damping=5e-4, limits=(-0.6, 0.6),
                                     kp=0.3, kv=1e-3)
    contacts = [(link.name, '') for link in links]
    standing = 0.8*leg_len + r_leg + 0.002
    spec = _finish(
        'centipede', links, joints_cfg, swimming=False,
        drag_coefficients=[[0.0]*3, [0.0]*3],
        water=WaterOptions(height=None), arena_z=0.0,
        spawn_pose=[0.0, 0.0, standing, 0.0, 0.0, 0.0], contacts_names=contacts,
        timestep=timestep, n_iterations=n_iterations, friction=[1.0, 0.0, 0.0])
    return spec


MODELS = {
    'swimmer8': swimmer8,
    'salamander': salamander,
    'salamander_swim': lambda **kw: salamander(swimming=True, **kw),
    'centipede': centipede,
}


def travelling_wave_parameters(spec, amplitude=0.3, frequency=1.0, wavenumber=1.0):
    """Travelling-wave position-control table for the body joints.

    ``ctrl[actuator_position_<j>] = A*sin(2*pi*(f*t - k*i/n) + phase_env)`` for
    the i-th body joint; leg joints are held at 0.
    Returns (joint_names, amplitude[], frequency[], phase_lag[]).
    """
    body = [j for j in spec.joints_names if 'leg' not in j]
    n = max(1, len(body))
    amp, freq, lag = [], [], []
    for jname in spec.joints_names:
        if jname in body:
            amp.append(amplitude)
            freq.append(frequency)
            lag.append(2.0*np.pi*wavenumber*body.index(jname)/n)
        else:
            amp.append(0.0)
            freq.append(frequency)
            lag.append(0.0)
    return list(spec.joints_names), np.array(amp), np.array(freq), np.array(lag)
