"""SDF -> MJCF for primitive-geometry animats.

The reference builds its MJCF from the animat's SDF file through farms_core's SDF reader and
dm_control (``setup_mjcf_xml`` farms_mujoco/simulation/mjcf.py:1174-1512, ``sdf2mjcf``
mjcf.py:132-600, called by ``Simulation.from_sdf`` simulation.py:96-124); neither exists in this
image.  This module covers the subset of that conversion the stepping path can run: links with
``<inertial>`` and sphere / capsule / box / cylinder / ellipsoid ``<collision>`` geometry, revolute
/ prismatic (and fixed) joints forming a tree; link, inertial and joint frames may be rotated (the
joint axis is read in the joint frame, which sits in the child link's, SDF 1.5+).  Meshes,
heightmaps, ball / universal joints and limitless joints raise ``NotImplementedError`` naming the
element.
The arena is the flat ground plane (and the water surface) of ``arena_options``; its own SDF is not
read.  The MJCF text follows the reference's schema and naming rules through the same emitter as
the synthetic models (models.py).
"""

import os
import xml.etree.ElementTree as ET

import numpy as np

from .mjcf_subset import euler_xyz2quat, quat2mat, quat_mul
from .models import AnimatSpec, _Geom, _Link, _emit_mjcf
from .options import JointOptions, LinkOptions


def _pose(element):
    node = element.find('pose') if element is not None else None
    if node is None or not (node.text or '').strip():
        return np.zeros(6)
    values = np.array([float(v) for v in node.text.split()])
    assert values.size == 6, f'<pose> needs six numbers: {node.text!r}'
    return values


def _number(element, tag, default=None):
    node = element.find(tag)
    if node is None:
        if default is None:
            raise ValueError(f'<{element.tag}> lacks <{tag}>')
        return default
    return float(node.text)


def _geometry(collision, name):
    geometry = collision.find('geometry')
    assert geometry is not None, f'collision {name}: no <geometry>'
    pose = _pose(collision)
    shape = list(geometry)[0]
    if shape.tag == 'sphere':
        size = (_number(shape, 'radius'), 0.0, 0.0)
    elif shape.tag in ('capsule', 'cylinder'):
        size = (_number(shape, 'radius'), 0.5*_number(shape, 'length'), 0.0)     # axis: local z, as in MuJoCo
    elif shape.tag == 'box':
        size = tuple(0.5*float(v) for v in shape.find('size').text.split())
    elif shape.tag == 'ellipsoid':
        size = tuple(float(v) for v in shape.find('radii').text.split())
    else:
        raise NotImplementedError(f'collision {name}: <{shape.tag}> geometry (primitive shapes only)')
    return _Geom(name=name, type=shape.tag, size=size, pos=tuple(pose[:3]), quat=tuple(euler_xyz2quat(pose[3:])))


def read_sdf(source):
    """``(model name, [_Link ...] in depth-first order, base link first)`` of the first model of an
    SDF file (path) or SDF text."""
    text = source
    if '<' not in source:
        with open(os.path.expandvars(source), encoding='utf-8') as sdf_file:
            text = sdf_file.read()
    root = ET.fromstring(text)
    model = root if root.tag == 'model' else root.find('model')
    assert model is not None, 'no <model> in the SDF'
    poses, links = {}, {}
    for node in model.findall('link'):
        name = node.get('name')
        pose = _pose(node)
        poses[name] = (pose[:3], np.asarray(euler_xyz2quat(pose[3:]), dtype=float))
        inertial = node.find('inertial')
        mass, ipos, diag, off = 0.0, np.zeros(3), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)
        if inertial is not None:
            ipose = _pose(inertial)
            mass, ipos = _number(inertial, 'mass'), ipose[:3]
            inertia = inertial.find('inertia')
            if inertia is not None:
                ixx, iyy, izz, ixy, ixz, iyz = (_number(inertia, k, 0.0) for k in ('ixx', 'iyy', 'izz', 'ixy', 'ixz', 'iyz'))
                # the tensor is given in the inertial frame: express it in the link's axes
                rot = quat2mat(np.asarray(euler_xyz2quat(ipose[3:]), dtype=float))
                tensor = rot @ np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]]) @ rot.T
                diag = (tensor[0, 0], tensor[1, 1], tensor[2, 2])
                off = (tensor[0, 1], tensor[0, 2], tensor[1, 2])
        if node.find('visual/geometry/mesh') is not None and node.find('collision') is None:
            raise NotImplementedError(f'link {name}: mesh-only link (no primitive <collision>)')
        geoms = [_geometry(collision, collision.get('name') or f'{name}_collision_{i}')
                 for i, collision in enumerate(node.findall('collision'))]
        links[name] = _Link(name=name, parent='', pos=(0.0, 0.0, 0.0), mass=mass, ipos=tuple(ipos),
                            inertia=diag, offdiag=off, geoms=geoms)
    children = {}
    is_child = set()
    for node in model.findall('joint'):
        name, kind = node.get('name'), node.get('type')
        parent, child = node.find('parent').text.strip(), node.find('child').text.strip()
        assert parent in links and child in links, f'joint {name}: unknown link'
        assert child not in is_child, f'joint {name}: link {child} has two parents (not a tree)'
        is_child.add(child)
        children.setdefault(parent, []).append(child)
        link = links[child]
        link.parent = parent
        # the child's frame in the parent's: both <pose>s are given in the model frame
        (p_pos, p_quat), (c_pos, c_quat) = poses[parent], poses[child]
        p_conj = p_quat*np.array([1.0, -1.0, -1.0, -1.0])
        link.pos = tuple(quat2mat(p_quat).T @ (c_pos - p_pos))
        link.quat = tuple(quat_mul(p_conj, c_quat))
        if kind == 'fixed':
            continue
        if kind not in ('revolute', 'continuous', 'prismatic'):
            raise NotImplementedError(f'joint {name}: type {kind!r} (revolute, prismatic and fixed joints only)')
        jpose = _pose(node)
        axis = node.find('axis')
        link.joint = name
        link.jtype = 'slide' if kind == 'prismatic' else 'hinge'
        link.jpos = tuple(jpose[:3])
        # <xyz> is given in the joint frame, which sits at <pose> in the child link's frame
        jrot = quat2mat(np.asarray(euler_xyz2quat(jpose[3:]), dtype=float))
        link.axis = tuple(jrot @ np.array([float(v) for v in axis.find('xyz').text.split()]))
        limit = axis.find('limit')
        link.limits = ((_number(limit, 'lower'), _number(limit, 'upper'))
                       if limit is not None and kind != 'continuous' else None)
    bases = [name for name in links if name not in is_child]
    assert len(bases) == 1, f'one base link expected, found {bases}'
    ordered = []

    def visit(name):
        ordered.append(links[name])
        for child in children.get(name, []):
            visit(child)

    visit(bases[0])
    assert len(ordered) == len(links)
    # the base link's own rotation (its frame in the model's)
    ordered[0].pos, ordered[0].quat = tuple(poses[bases[0]][0]), tuple(poses[bases[0]][1])
    return model.get('name') or 'animat', ordered


def spec_from_sdf(simulation_options, animat_options, arena_options, contacts_names=None):
    """``AnimatSpec`` (MJCF text + options) of ``animat_options.sdf``: what ``setup_mjcf_xml``
    (mjcf.py:1174-1512) hands to ``Simulation.__init__`` for a primitive-geometry animat on the
    flat arena.  Joint properties, motors, link options and the spawn pose come from the options,
    as in the reference (mjcf.py:647-866); joints the options do not list keep the SDF's limits and
    are passive."""
    model_name, links = read_sdf(animat_options.sdf)
    name = animat_options.name if animat_options.name != 'animat' else model_name
    listed = {joint.name: joint for joint in animat_options.morphology.joints}
    joints_opts = {}
    for link in links:
        if not link.joint:
            continue
        options = listed.get(link.joint) or JointOptions(name=link.joint)
        if options.limits:
            link.limits = tuple(options.limits)
        if link.limits is None:
            raise NotImplementedError(f'joint {link.joint}: no limits (the path handles limited hinges)')
        joints_opts[link.joint] = options
    link_opts = {link.name: link for link in animat_options.morphology.links}
    frictions = [tuple(link_opts[link.name].friction) for link in links if link.name in link_opts]
    friction = frictions[0] if frictions else (1.0, 0.0, 0.0)
    if any(f != friction for f in frictions):
        raise NotImplementedError('per-link friction (one friction triple for the animat)')
    water = arena_options.water
    mjcf, joint_order = _emit_mjcf(
        model_name=name, links=links, joints_opts=joints_opts, motors=animat_options.control.motors,
        spawn_pose=list(animat_options.spawn.pose), sim=simulation_options,
        arena_z=arena_options.ground_height if arena_options.ground_height is not None else 0.0,
        water_height=water.height, friction=friction,
        self_collisions=[tuple(pair) for pair in animat_options.morphology.self_collisions])
    if not animat_options.morphology.links:
        animat_options.morphology.links = [LinkOptions(name=link.name) for link in links]
    animat_options.morphology.joints = [joints_opts[j] for j in joint_order]
    swimming = [link.name for link in links if link.name in link_opts and link_opts[link.name].swimming]
    if contacts_names is None:
        contacts_names = [(link.name, '') for link in links if link.geoms]
    return AnimatSpec(
        name=name, mjcf=mjcf, animat_options=animat_options, arena_options=arena_options,
        simulation_options=simulation_options, links_names=[link.name for link in links],
        joints_names=joint_order, contacts_names=list(contacts_names), xfrc_names=swimming,
        base_link=links[0].name)
