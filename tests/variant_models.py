"""Hand-edited variants of the SWIMMER8 MJCF that leave the synthetic models' comfort zone:
slide joint, joint anchors off the body origin, tilted axes, joint stiffness + springref,
rotated body frame, general (non-axisymmetric) inertia with off-diagonals, ctrl / force
clamps, a geared motor, and a fixed base.  Used by the emulation and GPU parity tests."""

import copy
import dataclasses
import re

from farms_mujoco_b200 import models


def swimmer8_features():
    spec = models.swimmer8()
    x = spec.mjcf
    edits = [
        ('<joint name="joint_1" type="hinge" axis="0.0 0.0 1.0" pos="0.0 0.0 0.0"',
         '<joint name="joint_1" type="slide" axis="0.0 1.0 0.0" pos="0.01 0.0 0.0"'),
        ('<joint name="joint_2" type="hinge" axis="0.0 0.0 1.0" pos="0.0 0.0 0.0"',
         '<joint name="joint_2" type="hinge" axis="0.0 0.6 0.8" pos="0.02 0.01 -0.01"'),
        ('<joint name="joint_3" type="hinge" axis="0.0 0.0 1.0" pos="0.0 0.0 0.0" damping="0.001" '
         'stiffness="0.0" springref="0.0"',
         '<joint name="joint_3" type="hinge" axis="0.0 0.0 1.0" pos="0.0 0.0 0.0" damping="0.002" '
         'stiffness="0.05" springref="0.2"'),
        ('<body name="link_4" pos="0.1 0.0 0.0" quat="1.0 0.0 0.0 0.0">',
         '<body name="link_4" pos="0.1 0.0 0.0" quat="0.9689124217106447 0.1 0.2 0.1">'),
        ('<position name="actuator_position_joint_4" joint="joint_4" kp="1.0" ctrllimited="false" '
         'ctrlrange="-1000000.0 1000000.0"',
         '<position name="actuator_position_joint_4" joint="joint_4" kp="1.0" ctrllimited="true" '
         'ctrlrange="-0.05 0.05"'),
        ('<motor name="actuator_torque_joint_5" joint="joint_5"/>',
         '<motor name="actuator_torque_joint_5" joint="joint_5" forcelimited="true" '
         'forcerange="-0.01 0.01" gear="2.0"/>'),
    ]
    for old, new in edits:
        assert old in x, old
        x = x.replace(old, new)
    count = [0]

    def inertial(match):
        count[0] += 1
        if count[0] == 6:
            return ('<inertial pos="0.05 0.01 0.0" mass="0.2" '
                    'fullinertia="4e-05 0.00021 0.00025 1e-05 -2e-05 1.5e-05"/>')
        return match.group(0)

    x = re.sub(r'<inertial [^>]*/>', inertial, x)
    return dataclasses.replace(spec, name='swimmer8_features', mjcf=x)


def swimmer8_fixed_base():
    """farms 'fixed_base' (mjcf.py:751-756): no free joint; the base links fuse into the world."""
    return _fixed_base(models.swimmer8(), 'swimmer8_fixed')


def salamander_swim_fixed_base():
    """The same for a BRANCHING tree (legs and tail hang off a trunk whose first link is welded to
    the world): the tree-split kernels without a floating root."""
    return _fixed_base(models.salamander(swimming=True), 'salamander_swim_fixed')


def _fixed_base(spec, name):
    x = spec.mjcf
    root = re.search(r' *<freejoint name="[^"]*"/>\n', x)
    assert root, 'no free joint'
    x = x.replace(root.group(0), '')

    def trim(match):
        values = match.group(2).split()
        drop = 7 if match.group(1) == 'qpos' else 6
        return f'{match.group(1)}="{" ".join(values[drop:])}"'

    x = re.sub(r'(qpos|qvel)="([^"]*)"', trim, x)
    links = spec.links_names[1:]        # link_0 is welded to the world now: not a moving body
    animat = copy.deepcopy(spec.animat_options)
    animat.morphology.links = [link for link in animat.morphology.links if link.name in links]
    return dataclasses.replace(spec, name=name, mjcf=x, links_names=links,
                               animat_options=animat,
                               xfrc_names=links, contacts_names=[c for c in spec.contacts_names
                                                                 if c[0] in links])


def salamander_box_feet():
    """SALAMANDER on the ground with box feet (rotated about z by 0.3 rad) and a box trunk segment:
    plane-box contacts, the eight corners of mjc_PlaneBox."""
    spec = models.salamander()
    x = spec.mjcf
    n_feet = x.count('type="sphere" size="0.01 0.01 0.01" pos="0.0 0.0 -0.04" quat="1.0 0.0 0.0 0.0" friction=')
    assert n_feet == 4, n_feet
    x = x.replace('type="sphere" size="0.01 0.01 0.01" pos="0.0 0.0 -0.04" quat="1.0 0.0 0.0 0.0"',
                  'type="box" size="0.014 0.009 0.01" pos="0.0 0.0 -0.04" '
                  'quat="0.9887710779360422 0.0 0.0 0.14943813247359922"')
    old = ('type="capsule" size="0.020000000000000004 0.04 0.0" pos="0.04 0.0 0.0" '
           'quat="0.7071067811865476 0.0 0.7071067811865475 0.0"')
    assert x.count(old) == 2, x.count(old)
    x = x.replace(old, 'type="box" size="0.05 0.02 0.018" pos="0.04 0.0 0.0" quat="1.0 0.0 0.0 0.0"')
    return dataclasses.replace(spec, name='salamander_box', mjcf=x)


def salamander_ellipsoid_feet():
    """SALAMANDER on the ground with ellipsoid feet (radii 16 x 8 x 10 mm, rotated about z and tilted
    about x) and an ellipsoid trunk segment: plane-ellipsoid contacts, one contact at the support
    point along the plane normal (mjc_PlaneConvex)."""
    spec = models.salamander()
    x = spec.mjcf
    foot = 'type="sphere" size="0.01 0.01 0.01" pos="0.0 0.0 -0.04" quat="1.0 0.0 0.0 0.0"'
    assert x.count(foot) == 8, x.count(foot)       # collision geom + visual twin of four feet
    x = x.replace(foot, 'type="ellipsoid" size="0.016 0.008 0.01" pos="0.0 0.0 -0.04" '
                        'quat="0.9736691213452591 0.1471556949711443 0.02601839493044343 0.17215309088585662"')
    old = ('type="capsule" size="0.020000000000000004 0.04 0.0" pos="0.04 0.0 0.0" '
           'quat="0.7071067811865476 0.0 0.7071067811865475 0.0"')
    assert x.count(old) == 2, x.count(old)
    x = x.replace(old, 'type="ellipsoid" size="0.05 0.02 0.018" pos="0.04 0.0 0.0" quat="1.0 0.0 0.0 0.0"')
    return dataclasses.replace(spec, name='salamander_ellipsoid', mjcf=x)


def salamander_cylinder_feet():
    """SALAMANDER on the ground with cylinder feet (radius 12 mm, half length 6 mm, tilted) and a
    cylinder trunk segment lying along the body: plane-cylinder contacts, up to four points per
    cylinder (mjc_PlaneCylinder)."""
    spec = models.salamander()
    x = spec.mjcf
    foot = 'type="sphere" size="0.01 0.01 0.01" pos="0.0 0.0 -0.04" quat="1.0 0.0 0.0 0.0"'
    assert x.count(foot) == 8, x.count(foot)
    x = x.replace(foot, 'type="cylinder" size="0.012 0.006 0.0" pos="0.0 0.0 -0.04" '
                        'quat="0.9736691213452591 0.1471556949711443 0.02601839493044343 0.17215309088585662"')
    old = ('type="capsule" size="0.020000000000000004 0.04 0.0" pos="0.04 0.0 0.0" '
           'quat="0.7071067811865476 0.0 0.7071067811865475 0.0"')
    assert x.count(old) == 2, x.count(old)
    x = x.replace(old, 'type="cylinder" size="0.02 0.04 0.0" pos="0.04 0.0 0.0" '
                       'quat="0.7071067811865476 0.0 0.7071067811865475 0.0"')
    return dataclasses.replace(spec, name='salamander_cylinder', mjcf=x)


def salamander_foot_pairs(swimming=True, foot_radius=0.017, friction=0.5):
    """Explicit ``<contact><pair>`` self-collisions as ``mjcf.py:1012-1033`` writes them (condim 3,
    friction 0 -> MuJoCo's floor 1e-5, solref of the animat) between the left and the right foot
    of each girdle, with feet large enough to meet when the legs fold under the trunk; contact
    sensors between the two links and on each foot alone.  The rows of such a contact live on two
    branches of the tree."""
    spec = models.salamander(swimming=swimming)
    x = spec.mjcf
    old = 'size="0.01 0.01 0.01"'
    assert x.count(old) == 8, x.count(old)      # the four feet and their visual twins
    x = x.replace(old, f'size="{foot_radius} {foot_radius} {foot_radius}"')
    pairs = ''.join(
        f'    <pair name="contact_pair_{i}_0_0" geom1="link_leg_{i}_L_3_foot" geom2="link_leg_{i}_R_3_foot" '
        f'condim="3" friction="{friction} {friction} 0 0 0" solref="0.005 1"/>\n' for i in range(2))
    assert '<contact>' not in x
    x = x.replace('</mujoco>', f'  <contact>\n{pairs}  </contact>\n</mujoco>')
    contacts = list(spec.contacts_names) + [(f'link_leg_{i}_L_3', f'link_leg_{i}_R_3') for i in range(2)]
    return dataclasses.replace(spec, name='salamander_foot_pairs', mjcf=x, contacts_names=contacts)


def folded_legs_qpos(model, fold=1.1):
    """Keyframe with the four legs folded under the trunk (hip pitch and elbow at -/+ ``fold``):
    the feet of one girdle overlap."""
    import numpy as np
    qpos = np.array(model.key_qpos, dtype=float)
    for i in range(2):
        for side, sgn in (('L', 1.0), ('R', -1.0)):
            for k in (1, 3):
                adr = model.jnt_qposadr[model.jnt_id(f'joint_leg_{i}_{side}_{k}')]
                qpos[adr] = -sgn*fold
    return qpos
