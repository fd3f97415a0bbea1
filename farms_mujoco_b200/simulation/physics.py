"""Index maps and farms tables for the fused log stage.

Mirrors the init-time half of farms_mujoco/simulation/physics.py: the
name -> index maps of ``get_sensor_maps`` (physics.py:64-185) and
``get_physics2data_maps`` (physics.py:188-393), with the same dictionary keys
and the same integer contents.  The per-step half of
that file (``physics2data`` and its helpers, physics.py:423-545) is not host
code here: the gathers run inside the CUDA step kernel, driven by the
``FarmsTables`` compiled from these maps.

Instead of dm_control's ``physics.named`` indexers the functions take the flat
``Model`` (mjcf_subset.py), whose ``*_names`` lists play the role of the
``axes.row.names`` of the reference.
"""

import numpy as np

from ..mjcf_subset import JNT_FREE

# Sensor-name prefixes scanned by the reference, in its order (physics.py:70-85)
SENSOR_PREFIXES = (
    'framepos', 'framequat', 'framelinvel', 'frameangvel',
    'jointpos', 'jointvel', 'jointlimitfrc',
    'force', 'torque',
    'actuatorfrc_position', 'actuatorfrc_velocity', 'actuatorfrc_torque',
    'musclefrc', 'tendonpos', 'tendonvel',
    'musclefiberlen', 'musclefibervel',
    'musclepenn', 'muscleactivefrc', 'musclepassivefrc',
    'muscleIa', 'muscleII', 'muscleIb',
    'touch',
)


def get_sensor_maps(model, verbose=False):
    """Per-prefix ``{'names': [...], 'indices': int[n, dim]}`` (physics.py:64-110).

    A sensor belongs to a prefix when its name *starts with* it, so
    ``torque`` never matches ``actuatorfrc_torque_*`` but ``force`` would match
    a ``force_<joint>`` site sensor.
    """
    del verbose  # the reference only logs here (physics.py:112-185)
    maps = {}
    for prefix in SENSOR_PREFIXES:
        names, indices = [], []
        for sid, name in enumerate(model.sensor_names):
            if name.startswith(prefix):
                adr, dim = int(model.sensor_adr[sid]), int(model.sensor_dim[sid])
                names.append(name)
                indices.append(np.arange(adr, adr + dim))
        maps[prefix] = {'names': names, 'indices': np.array(indices, dtype=int)}
    return maps


def _first_address(model, kind, name):
    """``row2index(..., single=True)``: first qpos / dof address of a joint."""
    jid = model.jnt_id(name)
    return int(model.jnt_qposadr[jid] if kind == 'qpos' else model.jnt_dofadr[jid])


def get_physics2data_maps(model, sensor_data, sensor_maps):
    """Fill ``sensor_maps`` with the ``*2data`` index arrays (physics.py:188-393)."""
    links = list(sensor_data.links.names)
    joints = list(sensor_data.joints.names)
    body_rows = np.array([model.body_id(name) for name in links], dtype=int)
    for key in ('xpos2data', 'xquat2data', 'xipos2data', 'cvel2data'):
        sensor_maps[key] = body_rows.copy()
    sensor_maps['qpos2data'] = np.array([_first_address(model, 'qpos', j) for j in joints], dtype=int)
    sensor_maps['qvel2data'] = np.array([_first_address(model, 'qvel', j) for j in joints], dtype=int)

    # all-or-nothing per sensor family (physics.py:231-265)
    def family(prefix, items, first_only):
        have = sensor_maps[prefix]['names']
        wanted = [f'{prefix}_{item}' for item in items]
        if not all(name in have for name in wanted):
            return []
        rows = [sensor_maps[prefix]['indices'][have.index(name)] for name in wanted]
        return np.array([row[0] for row in rows] if first_only else rows, dtype=int)

    for prefix in ('framepos', 'framequat', 'framelinvel', 'frameangvel'):
        sensor_maps[f'{prefix}2data'] = family(prefix, links, first_only=False)
    if len(sensor_maps['framequat2data']) > 0:
        # MuJoCo wxyz -> farms xyzw (physics.py:243-246)
        sensor_maps['framequat2data'] = sensor_maps['framequat2data'][:, [1, 2, 3, 0]]
    for prefix in ('jointpos', 'jointvel', 'jointlimitfrc', 'actuatorfrc_position',
                   'actuatorfrc_velocity', 'actuatorfrc_torque'):
        sensor_maps[f'{prefix}2data'] = family(prefix, joints, first_only=True)
    for prefix in ('force', 'torque'):
        rows = [
            np.arange(model.sensor_adr[sid], model.sensor_adr[sid] + model.sensor_dim[sid])
            for j in joints
            for sid in [model.sensor_names.index(f'{prefix}_{j}')
                        if f'{prefix}_{j}' in model.sensor_names else -1]
            if sid >= 0
        ]
        sensor_maps[f'{prefix}2data'] = np.array(rows, dtype=int)

    # muscles are outside the hot path; keys kept so callers can test emptiness
    for prefix in ('musclefrc', 'musclefiberlen', 'musclefibervel', 'musclepenn',
                   'muscleactivefrc', 'musclepassivefrc', 'muscleIa', 'muscleII', 'muscleIb'):
        sensor_maps[f'{prefix}2data'] = []
    sensor_maps['tendonpos2data'] = np.array([], dtype=int)
    sensor_maps['tendonvel2data'] = np.array([], dtype=int)
    sensor_maps['musclesensors2data'] = np.zeros((0, 7), dtype=int)

    # actuator_moment is nu x nv; flat index = row*ncols + col (physics.py:344-357)
    nu, nv = model.nu, model.nv
    sensor_maps['actuator_moment2data'] = (
        np.arange(nu)[:, None]*nv + np.arange(nv)[None, :]).ravel()

    # contacts: (geom, -1) for "(body, '')" sensors, (geom1, geom2) for body pairs
    pairs = [tuple(pair) for pair in sensor_data.contacts.names]
    for pair in pairs:
        assert not isinstance(pair, str) and len(pair) == 2, (
            f'Contact "{pair}" should be a pair of strings')
    geom_body_names = [model.body_names[b] for b in model.geom_bodyid]
    geompair2data = {}
    for gid, bname in enumerate(geom_body_names):
        if (bname, '') in pairs:
            geompair2data[(gid, -1)] = pairs.index((bname, ''))
    for gid1, bname1 in enumerate(geom_body_names):
        for gid2, bname2 in enumerate(geom_body_names):
            if (bname1, bname2) in pairs:
                geompair2data[(gid1, gid2)] = pairs.index((bname1, bname2))
    found = set(geompair2data.values())
    for index, pair in enumerate(pairs):
        assert index in found, f'Missing pair: {pair} (body_names={model.body_names})'
    sensor_maps['geompair2data'] = geompair2data

    # external forces (physics.py:385-393)
    sensor_maps['data2xfrc'] = np.array(
        [model.body_id(name) for name in sensor_data.xfrc.names], dtype=int)
    sensor_maps['datalinks2xfrc'] = body_rows.copy()
    return sensor_maps


class FarmsTables:
    """Integer / real tables the device log stage and the fused drag consume.

    Built from the maps above plus the constructor arguments of the reference's
    ``SwimmingHandler`` (drag.pyx:333-387).  One instance is marshalled into
    ``FbFarms`` (include/farms_b200.h).
    """
    # pylint: disable=too-many-instance-attributes

    def __init__(self, model, sensor_data, sensor_maps, animat_options=None,
                 arena_options=None, units=None):
        # pylint: disable=too-many-locals,too-many-arguments
        from ..units import SimulationUnitScaling  # local: keep import graph flat
        units = units if units is not None else SimulationUnitScaling()
        joints = list(sensor_data.joints.names)
        self.link_body = np.asarray(sensor_maps['xpos2data'], dtype=np.int32)
        self.joint_qposadr = np.asarray(sensor_maps['qpos2data'], dtype=np.int32)
        self.joint_dofadr = np.asarray(sensor_maps['qvel2data'], dtype=np.int32)

        sensor_obj = {name: int(model.sensor_objid[i]) for i, name in enumerate(model.sensor_names)}

        def objects(key, prefix):
            if len(sensor_maps[key]) == 0:
                return -np.ones(len(joints), dtype=np.int32)
            return np.array([sensor_obj[f'{prefix}_{j}'] for j in joints], dtype=np.int32)

        self.joint_jntid = objects('jointlimitfrc2data', 'jointlimitfrc')
        self.joint_act_position = objects('actuatorfrc_position2data', 'actuatorfrc_position')
        self.joint_act_velocity = objects('actuatorfrc_velocity2data', 'actuatorfrc_velocity')
        # empty in the reference: the sensors are named actuatorfrc_motor_* 
        self.joint_act_torque = objects('actuatorfrc_torque2data', 'actuatorfrc_torque')

        self.n_contacts = len(sensor_data.contacts.names)
        geompair2data = sensor_maps['geompair2data']
        cand_sensor = -np.ones((model.ncand, 4), dtype=np.int32)
        for c in range(model.ncand):
            g1, g2 = int(model.cand_geom1[c]), int(model.cand_geom2[c])
            for k, key in enumerate(((g1, g2), (g2, g1), (g1, -1), (g2, -1))):
                cand_sensor[c, k] = geompair2data.get(key, -1)
        self.cand_sensor = cand_sensor
        self.xfrc_body = np.asarray(sensor_maps['data2xfrc'], dtype=np.int32)

        # swimming links (drag.pyx:353-385)
        swim = []
        if animat_options is not None and arena_options is not None \
                and arena_options.water.height is not None:
            swim = [link for link in animat_options.morphology.links if link.swimming]
        self.swim_links_index = np.array(
            [sensor_data.links.names.index(link.name) for link in swim], dtype=np.int32)
        self.swim_xfrc_index = np.array(
            [sensor_data.xfrc.names.index(link.name) for link in swim], dtype=np.int32)
        self.swim_mass = np.array(
            [model.body_mass[model.body_id(link.name)] for link in swim], dtype=float
        )/units.kilograms
        heights = []
        for link in swim:
            body = model.body_id(link.name)
            geoms = [g for g in range(model.ngeom) if model.geom_bodyid[g] == body]
            if not geoms:
                raise ValueError(f'swimming link {link.name} has no geom (drag.pyx:364-371)')
            heights.append(0.5*model.geom_rbound[geoms[0]])
        self.swim_height = np.array(heights, dtype=float)/units.meters
        self.swim_density = np.array([link.density for link in swim], dtype=float)
        self.swim_coefficients = np.array(
            [np.array(link.drag_coefficients, dtype=float) for link in swim], dtype=float
        ).reshape(len(swim), 2, 3)
        water = arena_options.water if arena_options is not None else None
        has_water = water is not None and water.height is not None
        self.water_drag = bool(has_water and water.drag)
        self.water_sph = bool(has_water and getattr(water, 'sph', False))
        self.water_buoyancy = bool(has_water and water.buoyancy)
        self.water_surface = float(water.height) if has_water else 0.0
        if self.water_sph:
            self.water_surface = 1e8  # drag.pyx:386-387
        self.water_density = float(water.density) if has_water else 1000.0
        self.water_viscosity = float(water.viscosity) if has_water else 1.0
        self.water_velocity = np.array(water.velocity if has_water else [0, 0, 0], dtype=float)
        self.meters, self.seconds, self.kilograms = units.as_tuple()
        # bodies whose root joint is free: used by host-side sanity checks only
        self.free_root = bool(model.njnt and model.jnt_type[0] == JNT_FREE)


def physics2data(physics, iteration, data, maps, units, links_only=False):
    """Sensors data collection (farms_mujoco/simulation/physics.py:527-545).

    The gathers, the unit scaling, the contact aggregation and the drag model of the
    reference run inside the CUDA step (csrc/fb_fast.h, csrc/fb_device.h: ring row
    ``iteration`` of the device log *is* what the reference writes into
    ``data.sensors.<kind>.array[iteration]``).  This mirror only copies that row to the
    host-side ``data`` (``BatchedAnimatData``: a leading environment axis), for callers
    that keep the reference's per-iteration protocol.  ``maps`` and ``units`` were
    compiled into the engine's tables at construction.
    """
    del maps, units
    sensors = data.sensors
    # With several physics steps per iteration (log_stride > 1) the refresh reads the state the
    # physics holds NOW, as the reference does: on a full step that is device row
    # iteration*log_stride, on a sub-step (links_only) it is a row between two iterations.
    latest = physics.log_stride > 1
    sensors.links.array[:, iteration] = physics.log_row('links', iteration, latest)
    if sensors.xfrc.array.shape[2]:
        # the drag forces of this row (drag.pyx:389-411 computes them from the links row that was
        # just refreshed, in the same before_step): the device wrote them with the row
        sensors.xfrc.array[:, iteration] = physics.log_row('xfrc', iteration, latest)
    if not links_only:
        sensors.joints.array[:, iteration] = physics.log_row('joints', iteration, latest)
        if sensors.contacts.array.shape[2]:
            sensors.contacts.array[:, iteration] = physics.log_row('contacts', iteration, latest)
