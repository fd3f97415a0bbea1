"""SDF front-end (sdf_subset.py; the primitive-geometry subset of the reference's
Simulation.from_sdf -> setup_mjcf_xml -> sdf2mjcf, simulation.py:96-124, mjcf.py:132-600)."""

import numpy as np
import pytest

from farms_mujoco_b200 import mjcf_subset, models, sdf_subset
from farms_mujoco_b200.options import (AnimatOptions, ArenaOptions, ControlOptions, JointOptions, LinkOptions,
                                       MorphologyOptions, MotorOptions, SimulationOptions, SpawnOptions, WaterOptions)


def swimmer8_sdf():
    """SWIMMER8 (models.swimmer8) written as an SDF model: absolute link poses, capsules along x."""
    radius, length = 0.02, 0.1
    mass, axial, transverse = models.capsule_mass_inertia(radius, length)
    out = ['<?xml version="1.0"?>', '<sdf version="1.6">', '  <model name="swimmer">']
    for i in range(8):
        out += [f'    <link name="link_{i}">', f'      <pose>{length*i!r} 0 0 0 0 0</pose>',
                '      <inertial>', f'        <pose>{0.5*length!r} 0 0 0 0 0</pose>', f'        <mass>{mass!r}</mass>',
                f'        <inertia><ixx>{axial!r}</ixx><iyy>{transverse!r}</iyy><izz>{transverse!r}</izz>'
                '<ixy>0</ixy><ixz>0</ixz><iyz>0</iyz></inertia>', '      </inertial>',
                f'      <collision name="link_{i}_collision">', f'        <pose>{0.5*length!r} 0 0 0 {0.5*np.pi!r} 0</pose>',
                f'        <geometry><capsule><radius>{radius!r}</radius><length>{length!r}</length></capsule></geometry>',
                '      </collision>', '    </link>']
    for i in range(7):
        out += [f'    <joint name="joint_{i}" type="revolute">', f'      <parent>link_{i}</parent>',
                f'      <child>link_{i + 1}</child>', '      <axis><xyz>0 0 1</xyz><limit><lower>-1.0</lower><upper>1.0</upper></limit></axis>',
                '    </joint>']
    out += ['  </model>', '</sdf>']
    return '\n'.join(out)


def swimmer8_options(sdf):
    reference = models.swimmer8()
    animat = AnimatOptions(
        name='swimmer', sdf=sdf, spawn=SpawnOptions(pose=list(reference.animat_options.spawn.pose)),
        morphology=MorphologyOptions(
            links=[LinkOptions(name=f'link_{i}', swimming=True, drag_coefficients=[[-0.1, -1.0, -1.0], [-1e-3]*3],
                               friction=[1.0, 0.0, 0.0]) for i in range(8)],
            joints=[JointOptions(name=f'joint_{i}', damping=1e-3, limits=[-1.0, 1.0]) for i in range(7)]),
        control=ControlOptions(motors=[MotorOptions(joint_name=f'joint_{i}', gains=[1.0, 1e-3]) for i in range(7)]))
    arena = ArenaOptions(ground_height=-2.0, water=WaterOptions(height=0.0, drag=True, buoyancy=True))
    return reference, SimulationOptions(timestep=1e-3, n_iterations=reference.simulation_options.n_iterations), animat, arena


def test_swimmer8_from_sdf_is_the_synthetic_model(tmp_path):
    """The SDF description of SWIMMER8 yields the same compiled model as the generator of the
    benchmark workload -- from text and from a file."""
    path = tmp_path/'swimmer.sdf'
    path.write_text(swimmer8_sdf())
    for source in (swimmer8_sdf(), str(path)):
        reference, sim, animat, arena = swimmer8_options(source)
        spec = sdf_subset.spec_from_sdf(sim, animat, arena)
        ours, ref = mjcf_subset.parse_mjcf(spec.mjcf), mjcf_subset.parse_mjcf(reference.mjcf)
        assert spec.links_names == reference.links_names and spec.joints_names == reference.joints_names
        assert spec.xfrc_names == reference.xfrc_names and spec.contacts_names == reference.contacts_names
        for field in ('body_parentid', 'body_pos', 'body_mass', 'body_ipos', 'body_inertia', 'jnt_axis', 'jnt_range',
                      'jnt_pos', 'geom_size', 'geom_pos', 'geom_quat', 'geom_type', 'key_qpos', 'actuator_gainprm',
                      'actuator_biasprm', 'dof_damping'):
            a, b = np.asarray(getattr(ours, field), dtype=float), np.asarray(getattr(ref, field), dtype=float)
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-15), field


def branching_sdf(extra='', leg_rpy=(0.0, 0.0, 0.0)):
    """A trunk of two links with a sphere-footed leg on each side, the second joint anchored off its
    link origin, and a fixed sensor link.  ``leg_rpy`` turns the FRAME of the left leg (not the leg):
    its inertial offset, inertia tensor, collision offset, joint anchor and joint axis are rewritten
    in the turned frame, so the animat is physically the same."""
    from farms_mujoco_b200.mjcf_subset import euler_xyz2quat, quat2mat
    rot = quat2mat(np.asarray(euler_xyz2quat(leg_rpy), dtype=float))          # leg frame -> model frame
    inertia = rot.T @ np.array([[4e-6, 1e-7, 0.0], [1e-7, 4e-6, 0.0], [0.0, 0.0, 2e-6]]) @ rot
    vec = lambda v: ' '.join(repr(float(x)) for x in rot.T @ np.asarray(v, dtype=float))
    leg = dict(rpy=' '.join(repr(float(a)) for a in leg_rpy), com=vec([0, 0, -0.02]), foot=vec([0, 0, -0.04]),
               anchor=vec([0, -0.01, 0]), axis=vec([0, 1, 0]),
               **{k: repr(float(inertia[i, j])) for k, (i, j) in dict(ixx=(0, 0), iyy=(1, 1), izz=(2, 2), ixy=(0, 1),
                                                                        ixz=(0, 2), iyz=(1, 2)).items()})
    return f"""<sdf version="1.6"><model name="walker">
  <link name="trunk_0"><pose>0 0 0 0 0 0</pose>
    <inertial><pose>0.05 0 0 0 0 0</pose><mass>0.2</mass><inertia><ixx>1e-4</ixx><iyy>3e-4</iyy><izz>3e-4</izz></inertia></inertial>
    <collision name="trunk_0_c"><pose>0.05 0 0 0 1.5707963267948966 0</pose><geometry><capsule><radius>0.02</radius><length>0.1</length></capsule></geometry></collision></link>
  <link name="trunk_1"><pose>0.1 0 0 0 0 0</pose>
    <inertial><pose>0.05 0 0 0 0 0</pose><mass>0.2</mass><inertia><ixx>1e-4</ixx><iyy>3e-4</iyy><izz>3e-4</izz></inertia></inertial>
    <collision name="trunk_1_c"><pose>0.05 0 0 0 0 0</pose><geometry><box><size>0.1 0.04 0.03</size></box></geometry></collision></link>
  <link name="leg_L"><pose>0.05 0.04 0 {leg['rpy']}</pose>
    <inertial><pose>{leg['com']} 0 0 0</pose><mass>0.02</mass><inertia><ixx>{leg['ixx']}</ixx><iyy>{leg['iyy']}</iyy><izz>{leg['izz']}</izz><ixy>{leg['ixy']}</ixy><ixz>{leg['ixz']}</ixz><iyz>{leg['iyz']}</iyz></inertia></inertial>
    <collision name="foot_L"><pose>{leg['foot']} 0 0 0</pose><geometry><sphere><radius>0.01</radius></sphere></geometry></collision></link>
  <link name="leg_R"><pose>0.05 -0.04 0 0 0 0</pose>
    <inertial><pose>0 0 -0.02 0 0 0</pose><mass>0.02</mass><inertia><ixx>4e-6</ixx><iyy>4e-6</iyy><izz>2e-6</izz></inertia></inertial>
    <collision name="foot_R"><pose>0 0 -0.04 0 0 0</pose><geometry><sphere><radius>0.01</radius></sphere></geometry></collision></link>
  <link name="imu"><pose>0.02 0 0.02 0 0 0</pose><inertial><mass>0.001</mass><inertia><ixx>1e-9</ixx><iyy>1e-9</iyy><izz>1e-9</izz></inertia></inertial></link>
  <joint name="spine" type="revolute"><parent>trunk_0</parent><child>trunk_1</child>
    <axis><xyz>0 0 1</xyz><limit><lower>-0.8</lower><upper>0.8</upper></limit></axis></joint>
  <joint name="hip_L" type="revolute"><parent>trunk_0</parent><child>leg_L</child><pose>{leg['anchor']} 0 0 0</pose>
    <axis><xyz>{leg['axis']}</xyz><limit><lower>-0.5</lower><upper>0.5</upper></limit></axis></joint>
  <joint name="hip_R" type="revolute"><parent>trunk_0</parent><child>leg_R</child>
    <axis><xyz>0 1 0</xyz><limit><lower>-0.5</lower><upper>0.5</upper></limit></axis></joint>
  <joint name="imu_mount" type="fixed"><parent>trunk_0</parent><child>imu</child></joint>
  {extra}
</model></sdf>"""


def test_branching_sdf():
    name, links = sdf_subset.read_sdf(branching_sdf())
    assert name == 'walker'
    assert [link.name for link in links] == ['trunk_0', 'trunk_1', 'leg_L', 'leg_R', 'imu']      # depth first, joint order
    by_name = {link.name: link for link in links}
    assert by_name['leg_L'].parent == 'trunk_0' and np.allclose(by_name['leg_L'].pos, [0.05, 0.04, 0.0])
    assert np.allclose(by_name['leg_L'].jpos, (0.0, -0.01, 0.0)) and np.allclose(by_name['leg_L'].offdiag, (1e-7, 0.0, 0.0))
    assert by_name['imu'].joint == '' and by_name['trunk_1'].geoms[0].size == (0.05, 0.02, 0.015)
    animat = AnimatOptions(sdf=branching_sdf(), spawn=SpawnOptions(pose=[0, 0, 0.05, 0, 0, 0]),
                           control=ControlOptions(motors=[MotorOptions(joint_name='spine', gains=[0.5, 1e-3])]))
    spec = sdf_subset.spec_from_sdf(SimulationOptions(), animat, ArenaOptions(ground_height=0.0))
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert spec.joints_names == ['spine', 'hip_L', 'hip_R'] and model.nu == 3 and model.nv == 9
    assert spec.contacts_names == [('trunk_0', ''), ('trunk_1', ''), ('leg_L', ''), ('leg_R', '')]
    assert spec.base_link == 'trunk_0' and [j.name for j in animat.morphology.joints] == spec.joints_names


@pytest.mark.parametrize('extra,message', [
    ('<link name="s"/><joint name="p" type="ball"><parent>trunk_0</parent><child>s</child></joint>', "'ball'"),
    ('<link name="h"><collision name="c"><geometry><mesh><uri>a.obj</uri></mesh></geometry></collision></link>', '<mesh>'),
])
def test_unsupported_elements_are_named(extra, message):
    with pytest.raises(NotImplementedError, match=message):
        sdf_subset.read_sdf(branching_sdf(extra))


def test_simulation_from_sdf_runs(emu_library):
    """Simulation.from_sdf end to end on the emulation build: the branching walker settles on its
    feet, the log has the reference's shapes."""
    from farms_mujoco_b200.simulation.simulation import Simulation
    animat = AnimatOptions(sdf=branching_sdf(), spawn=SpawnOptions(pose=[0, 0, 0.06, 0, 0, 0]),
                           control=ControlOptions(motors=[MotorOptions(joint_name=j, gains=[0.5, 1e-3])
                                                          for j in ('spine', 'hip_L', 'hip_R')]))
    sim = Simulation.from_sdf(SimulationOptions(timestep=1e-3, n_iterations=60), animat, ArenaOptions(ground_height=0.0),
                              n_envs=2, library=emu_library)
    sim.run()
    sensors = sim.task.data.sensors
    assert sensors.links.array.shape[-2:] == (5, 20) and sensors.joints.array.shape[-2:] == (3, 18)
    assert np.isfinite(sensors.links.array).all()
    assert np.abs(sensors.contacts.array[..., 2]).max() > 0.1          # the feet carry the weight (0.44 kg)


def test_rotated_link_frame_is_the_same_animat(emu_library):
    """Turning the FRAME of a link (with everything expressed in it rewritten accordingly) changes
    the MJCF -- body quaternion, joint axis, inertial -- but not the animat: world trajectories agree."""
    from farms_mujoco_b200.simulation.simulation import Simulation
    logs = []
    for rpy in ((0.0, 0.0, 0.0), (0.4, -0.3, 0.7)):
        animat = AnimatOptions(sdf=branching_sdf(leg_rpy=rpy), spawn=SpawnOptions(pose=[0, 0, 0.055, 0, 0, 0]),
                               control=ControlOptions(motors=[MotorOptions(joint_name=j, gains=[0.5, 1e-3])
                                                              for j in ('spine', 'hip_L', 'hip_R')]))
        sim = Simulation.from_sdf(SimulationOptions(timestep=1e-3, n_iterations=40), animat, ArenaOptions(ground_height=0.0),
                                  n_envs=1, library=emu_library)
        if any(rpy):
            leg = [line for line in sim._mjcf_model.splitlines() if '<body name="leg_L"' in line][0]
            assert 'quat="1.0 0.0 0.0 0.0"' not in leg
        sim.run()
        sensors = sim.task.data.sensors
        logs.append((sensors.links.array.copy(), sensors.joints.array.copy(), sensors.contacts.array.copy()))
    links0, joints0, contacts0 = logs[0]
    links1, joints1, contacts1 = logs[1]
    assert np.abs(links0[..., :3] - links1[..., :3]).max() < 2e-6          # CoM positions [m]
    assert np.abs(joints0[..., :2] - joints1[..., :2]).max() < 1e-4       # joint positions / velocities
    assert np.abs(contacts0[..., :3] - contacts1[..., :3]).max() < 2e-3*np.abs(contacts0[..., :3]).max()
    assert np.abs(contacts0[..., 2]).max() > 0.1


def test_prismatic_joint_and_rotated_inertial_and_joint_frames():
    """A prismatic joint becomes a slide joint; an <inertial> frame turned against the link's gives
    the tensor in link axes; <xyz> is read in the (turned) joint frame."""
    extra = """<link name="probe"><pose>0.1 0 0.03 0 0 0</pose>
      <inertial><pose>0 0 0 0 0 1.5707963267948966</pose><mass>0.01</mass><inertia><ixx>1e-6</ixx><iyy>3e-6</iyy><izz>2e-6</izz></inertia></inertial>
      <collision name="tip"><geometry><sphere><radius>0.005</radius></sphere></geometry></collision></link>
    <joint name="extend" type="prismatic"><parent>trunk_1</parent><child>probe</child><pose>0 0 0 0 1.5707963267948966 0</pose>
      <axis><xyz>0 0 1</xyz><limit><lower>0.0</lower><upper>0.02</upper></limit></axis></joint>"""
    _, links = sdf_subset.read_sdf(branching_sdf(extra))
    probe = [link for link in links if link.name == 'probe'][0]
    assert probe.jtype == 'slide' and probe.parent == 'trunk_1' and probe.limits == (0.0, 0.02)
    assert np.allclose(probe.inertia, (3e-6, 1e-6, 2e-6), atol=1e-18) and np.allclose(probe.offdiag, 0.0, atol=1e-18)   # x <-> y
    assert np.allclose(probe.axis, (1.0, 0.0, 0.0), atol=1e-12)                  # joint z turned onto the link's x
    animat = AnimatOptions(sdf=branching_sdf(extra), spawn=SpawnOptions(pose=[0, 0, 0.05, 0, 0, 0]))
    spec = sdf_subset.spec_from_sdf(SimulationOptions(), animat, ArenaOptions(ground_height=0.0))
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert 'type="slide"' in spec.mjcf and model.nv == 10 and list(model.jnt_type).count(2) == 1


def test_from_sdf_with_cylinder_and_ellipsoid_collisions(emu_library):
    """SDF <cylinder> and <ellipsoid> collisions reach the contact pipeline: a skid (cylinder lying
    across the trunk, welded on) and an ellipsoid tail pad carry part of the walker's weight."""
    from farms_mujoco_b200.simulation.simulation import Simulation
    extra = """<link name="skid"><pose>0.05 0 -0.045 0 0 0</pose><inertial><mass>0.005</mass><inertia><ixx>1e-7</ixx><iyy>1e-7</iyy><izz>1e-7</izz></inertia></inertial>
      <collision name="skid_c"><pose>0 0 0 1.5707963267948966 0 0</pose><geometry><cylinder><radius>0.008</radius><length>0.05</length></cylinder></geometry></collision></link>
    <joint name="skid_mount" type="fixed"><parent>trunk_0</parent><child>skid</child></joint>
    <link name="pad"><pose>0.2 0 -0.04 0 0 0</pose><inertial><mass>0.005</mass><inertia><ixx>1e-7</ixx><iyy>1e-7</iyy><izz>1e-7</izz></inertia></inertial>
      <collision name="pad_c"><pose>0 0 0 0 0.2 0</pose><geometry><ellipsoid><radii>0.02 0.01 0.012</radii></ellipsoid></geometry></collision></link>
    <joint name="pad_mount" type="fixed"><parent>trunk_1</parent><child>pad</child></joint>"""
    animat = AnimatOptions(sdf=branching_sdf(extra), spawn=SpawnOptions(pose=[0, 0, 0.06, 0, 0, 0]),
                           control=ControlOptions(motors=[MotorOptions(joint_name='spine', gains=[0.5, 1e-3])]))
    sim = Simulation.from_sdf(SimulationOptions(timestep=1e-3, n_iterations=80), animat, ArenaOptions(ground_height=0.0),
                              n_envs=2, library=emu_library)
    assert 'type="cylinder"' in sim._mjcf_model and 'type="ellipsoid"' in sim._mjcf_model
    sim.run()
    contacts = sim.task.data.sensors.contacts.array
    assert np.isfinite(contacts).all()
    # contact sensors in link order (depth first): trunk_0, trunk_1, pad, leg_L, leg_R, skid
    load = np.abs(contacts[..., 2]).max(axis=tuple(range(contacts.ndim - 2)))
    assert load.shape == (6,) and load[2] > 0.5 and load[5] > 0.5, load


def test_self_collisions_emit_the_reference_pairs(emu_library):
    """morphology.self_collisions -> <contact><pair> as mjcf.py:1012-1033 writes them (one per couple
    of collision geoms, condim 3, friction [0]*5).  Those frictionless pairs are refused with the
    reason (fp32 conditioning); with a friction coefficient the same pair runs."""
    from farms_mujoco_b200.options import MorphologyOptions
    from farms_mujoco_b200.simulation.simulation import Simulation
    animat = AnimatOptions(sdf=branching_sdf(), spawn=SpawnOptions(pose=[0, 0, 0.06, 0, 0, 0]),
                           morphology=MorphologyOptions(self_collisions=[['leg_L', 'leg_R']]),
                           control=ControlOptions(motors=[MotorOptions(joint_name=j, gains=[0.5, 1e-3])
                                                          for j in ('spine', 'hip_L', 'hip_R')]))
    spec = sdf_subset.spec_from_sdf(SimulationOptions(timestep=1e-3, n_iterations=20), animat,
                                    ArenaOptions(ground_height=0.0))
    line = '<pair name="contact_pair_0_0_0" geom1="foot_L" geom2="foot_R" condim="3" friction="0 0 0 0 0"/>'
    assert line in spec.mjcf
    with pytest.raises(NotImplementedError, match='friction'):
        mjcf_subset.parse_mjcf(spec.mjcf)
    import dataclasses
    spec = dataclasses.replace(spec, mjcf=spec.mjcf.replace('friction="0 0 0 0 0"', 'friction="0.6 0.6 0 0 0"'),
                               contacts_names=spec.contacts_names + [('leg_L', 'leg_R')])
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    assert (model.cand_end == 20).sum() == 1 and model.cand_friction[model.cand_end == 20][0] == 0.6
    sim = Simulation.from_spec(spec, n_envs=2, library=emu_library)
    assert sim.physics.fast_path == 0
    sim.run()
    assert np.isfinite(sim.task.data.sensors.links.array).all()
