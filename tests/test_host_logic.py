"""Host logic: MJCF front-end, index maps (bit-exact contract), layout, ABI."""

import ctypes as ct
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from farms_mujoco_b200 import mjcf_subset as ms
from farms_mujoco_b200 import models, cabi
from farms_mujoco_b200.data import AnimatData
from farms_mujoco_b200.layout import sc
from farms_mujoco_b200.simulation.physics import (
    FarmsTables, get_sensor_maps, get_physics2data_maps)


def _maps(spec):
    model = ms.parse_mjcf(spec.mjcf)
    data = AnimatData.from_sensors_names(1e-3, 2, spec.links_names, spec.joints_names,
                                         spec.contacts_names, spec.xfrc_names)
    maps = get_sensor_maps(model)
    get_physics2data_maps(model, data.sensors, maps)
    return model, data, maps


@pytest.mark.parametrize('name,nl,nj,nv,nu', [('swimmer8', 8, 7, 13, 21), ('salamander', 28, 27, 33, 81),
                                             ('centipede', 42, 41, 47, 123)])
def test_model_sizes_match_survey(name, nl, nj, nv, nu):
    spec = models.MODELS[name]()
    model = ms.parse_mjcf(spec.mjcf)
    assert (len(spec.links_names), len(spec.joints_names), model.nv, model.nu) == (nl, nj, nv, nu)
    assert model.nq == nv + 1
    assert model.body_names[0] == 'world' and 'arena' not in model.body_names   # fusestatic
    assert (np.diff(model.body_parentid[1:]) >= -model.nbody).all()
    assert (model.body_parentid[1:] < np.arange(1, model.nbody)).all()


def test_index_maps_contract():
    """Appendix E: keys, shapes and integer contents of the maps (physics.py:188-393)."""
    spec = models.salamander(swimming=True)
    model, data, maps = _maps(spec)
    nl, nj = len(spec.links_names), len(spec.joints_names)
    for key in ('xpos2data', 'xquat2data', 'xipos2data', 'cvel2data', 'datalinks2xfrc'):
        assert maps[key].tolist() == [model.body_id(n) for n in spec.links_names]
    assert maps['qpos2data'].tolist() == [7 + i for i in range(nj)]
    assert maps['qvel2data'].tolist() == [6 + i for i in range(nj)]
    assert maps['framelinvel2data'].shape == (nl, 3) and maps['frameangvel2data'].shape == (nl, 3)
    # the model-root body has sensors too, so link 0's sensors start after them
    assert maps['framelinvel2data'][0].tolist() == [6, 7, 8]
    assert maps['frameangvel2data'][0].tolist() == [9, 10, 11]
    assert len(maps['framepos2data']) == 0 and len(maps['framequat2data']) == 0
    for key in ('jointpos2data', 'jointvel2data', 'jointlimitfrc2data',
                'actuatorfrc_position2data', 'actuatorfrc_velocity2data'):
        assert maps[key].shape == (nj,)
    # reference quirk D-4: motor force sensors are named actuatorfrc_motor_*, so this map is empty
    assert len(maps['actuatorfrc_torque2data']) == 0
    assert maps['actuatorfrc_position']['names'][0] == f'actuatorfrc_position_{spec.joints_names[0]}'
    assert maps['data2xfrc'].tolist() == [model.body_id(n) for n in spec.xfrc_names]
    assert maps['actuator_moment2data'].shape == (model.nu*model.nv,)
    # contacts: every (link, '') sensor is reachable from each geom of that link
    gp = maps['geompair2data']
    for gid, body in enumerate(model.geom_bodyid):
        name = model.body_names[body]
        if (name, '') in spec.contacts_names:
            assert gp[(gid, -1)] == spec.contacts_names.index((name, ''))
        else:
            assert (gid, -1) not in gp
    foot_link = 'link_leg_0_L_3'
    geoms = [g for g in range(model.ngeom) if model.body_names[model.geom_bodyid[g]] == foot_link]
    assert len(geoms) == 2 and len({gp[(g, -1)] for g in geoms}) == 1


def test_missing_contact_pair_asserts():
    spec = models.swimmer8()
    spec.contacts_names.append(('no_such_link', ''))
    with pytest.raises(AssertionError, match='Missing pair'):
        _maps(spec)


def test_farms_tables_swimming_constants():
    """SwimmingHandler.__init__ outputs (drag.pyx:353-387)."""
    spec = models.swimmer8()
    model, data, maps = _maps(spec)
    tables = FarmsTables(model, data.sensors, maps, spec.animat_options, spec.arena_options,
                         spec.simulation_options.units)
    assert tables.swim_links_index.tolist() == list(range(8))
    assert np.allclose(tables.swim_mass, model.body_mass[2:10])
    # height = 0.5 * rbound of the first geom of the body; capsule rbound = r + halflen
    assert np.allclose(tables.swim_height, 0.5*(0.02 + 0.05))
    assert tables.swim_coefficients.shape == (8, 2, 3)
    assert tables.water_drag and tables.water_buoyancy and tables.water_surface == 0.0
    assert tables.cand_sensor.shape == (model.ncand, 4)
    assert (tables.cand_sensor[:, [0, 1, 2]] == -1).all()       # floor has no sensor; no pairs
    assert (tables.cand_sensor[:, 3] >= 0).all()
    spec.arena_options.water.sph = True
    tables = FarmsTables(model, data.sensors, maps, spec.animat_options, spec.arena_options)
    assert tables.water_surface == 1e8                           # drag.pyx:386-387


def test_layout_contract():
    assert (sc.link_size, sc.contact_size, sc.xfrc_size) == (20, 12, 6)
    assert sc.link_com_position_x == 0                 # drag.pyx:189-191 reads cols 0..2
    for base in ('link_urdf_position', 'link_com_position', 'link_com_velocity_lin', 'link_com_velocity_ang'):
        x, y, z = (getattr(sc, f'{base}_{a}') for a in 'xyz')
        assert (y, z) == (x + 1, x + 2)
    for base in ('link_urdf_orientation', 'link_com_orientation'):
        x, w = getattr(sc, f'{base}_x'), getattr(sc, f'{base}_w')
        assert w == x + 3
    cols = [getattr(sc, f'contact_{k}_{a}') for k in ('reaction', 'friction', 'total', 'position') for a in 'xyz']
    assert cols == list(range(12))


def test_mjcf_compiler_numbers():
    """fullinertia -> principal axes; invweight0; capsule rbound; candidate order."""
    spec = models.centipede()
    model = ms.parse_mjcf(spec.mjcf)
    leg = model.body_id('link_leg_3_L')
    rot = ms.quat2mat(model.body_iquat[leg])
    full = rot @ np.diag(model.body_inertia[leg]) @ rot.T
    xml_full = re.search(r'<body name="link_leg_3_L".*?fullinertia="([^"]+)"', spec.mjcf, re.S).group(1)
    fi = np.array([float(v) for v in xml_full.split()])
    assert np.allclose(full, [[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]], atol=1e-15)
    assert (np.diff(model.body_inertia[leg]) <= 1e-18).all()     # descending, like mju_eig3
    assert (model.dof_invweight0 > 0).all() and (model.body_invweight0[2:, 0] > 0).all()
    assert np.allclose(model.dof_invweight0[:3], model.dof_invweight0[0])
    assert model.ncand == 84 and model.cand_end[:2].tolist() == [1, -1]
    assert model.geom_rbound[model.geom_id('link_body_0_collision')] == pytest.approx(0.015 + 0.025)


def test_mjcf_rejects_out_of_scope():
    spec = models.swimmer8()
    with pytest.raises(NotImplementedError):
        ms.parse_mjcf(spec.mjcf.replace('cone="pyramidal"', 'cone="elliptic"'))
    with pytest.raises(ValueError, match='unknown geom'):
        ms.parse_mjcf(spec.mjcf.replace('</mujoco>', '<contact><pair geom1="a" geom2="b"/></contact></mujoco>'))
    with pytest.raises(NotImplementedError, match='sphere-sphere'):      # capsule pairs: not yet
        ms.parse_mjcf(spec.mjcf.replace('</mujoco>', '<contact><pair geom1="link_0_collision" '
                                        'geom2="link_2_collision"/></contact></mujoco>'))
    with pytest.raises(NotImplementedError, match='exclude'):
        ms.parse_mjcf(spec.mjcf.replace('</mujoco>', '<contact><exclude body1="link_0" body2="link_1"/></contact></mujoco>'))


def test_abi_header_symbols_are_exported(cuda_library):
    """Every function include/farms_b200.h declares is exported by the .so and bound
    by the ctypes layer (no compute calls: this runs without a GPU)."""
    header = open(os.path.join(ROOT, 'include', 'farms_b200.h')).read()
    declared = set(re.findall(r'\b(fb_[a-z_0-9]+)\s*\(', header))
    from farms_mujoco_b200 import engine
    assert declared == set(engine.ABI_SYMBOLS), declared ^ set(engine.ABI_SYMBOLS)
    lib = ct.CDLL(cuda_library)
    for name in declared:
        assert hasattr(lib, name), name
    assert engine.load_library(cuda_library).fb_abi_version() == 2


def test_struct_mirrors_match_header_field_order():
    header = open(os.path.join(ROOT, 'include', 'farms_b200.h')).read()
    for cname, cls in (('FbModel', cabi.FbModel), ('FbFarms', cabi.FbFarms),
                       ('FbWaveController', cabi.FbWaveController), ('FbLogView', cabi.FbLogView),
                       ('FbStateView', cabi.FbStateView), ('FbDerivedView', cabi.FbDerivedView)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r'^(const\s+)?(int32_t|int64_t|double|float)\s+', '', decl)
            for part in decl.split(','):
                names.append(re.sub(r'[\*\s]|\[\d+\]', '', part))
        assert names == [f[0] for f in cls._fields_], cname


def test_no_cpu_fallback_without_library(tmp_path):
    from farms_mujoco_b200 import engine
    with pytest.raises(engine.EngineError, match='no CPU fallback'):
        engine.load_library(str(tmp_path/'missing.so'))


def test_tree_split_schedule(emu_library):
    """FastSplit (fb_model.h: fb_build_split): the trunk up to the last branching body on warp 0, the
    subtrees behind it in parallel, subtrees of one parent on one warp, every body exactly once, and
    every chain link (parent == body - 1) consecutive on its warp."""
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    for name, want_warps in (('swimmer8', 1), ('salamander_swim', 3), ('centipede', 4)):
        spec = models.MODELS[name]()
        model = mjcf_subset.parse_mjcf(spec.mjcf)
        physics = BatchedPhysics.from_spec(spec, 2, buffer_size=2, library=emu_library)
        sched = physics.fast_split_schedule()
        assert len(sched) == want_warps, (name, sched)
        parent = [int(p) for p in model.body_parentid]
        seen = sorted(b for a, bb in sched for b in a + bb)
        assert seen == list(range(1, model.nbody)), name
        if want_warps == 1:
            continue
        trunk = sched[0][0]
        split = trunk[-1]
        assert trunk == list(range(1, split + 1)) and all(not a for a, _ in sched[1:])
        assert max(parent[b] for b in range(2, model.nbody) if parent[b] != b - 1 and parent[b] >= 1) == split
        owner = {b: w for w, (a, bb) in enumerate(sched) for b in a + bb}
        for w, (a, bb) in enumerate(sched):
            order = a + bb
            for i, b in enumerate(order):
                p = parent[b]
                if b > split and p > split:
                    assert owner[p] == w and order.index(p) < i            # inside its subtree
                if p == b - 1 and p >= 1:
                    assert i > 0 and order[i - 1] == p, (name, b)          # register hand-over stays consecutive
        roots = {}
        for b in range(split + 1, model.nbody):
            if parent[b] <= split and not (b == split + 1 and parent[b] == split):
                roots.setdefault(parent[b], set()).add(owner[b])
        assert all(len(ws) == 1 for ws in roots.values()), roots     # one slot, one warp
    if name == 'salamander_swim':
        pass
