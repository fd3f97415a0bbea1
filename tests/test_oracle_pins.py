"""Pins for the fp64 CPU oracle (PARITY UNPINNED against MuJoCo: the reference
ships no golden vectors and MuJoCo is absent).  These are
the substitute pins: analytic known answers and self-consistency identities."""

import numpy as np
import pytest

from farms_mujoco_b200 import mjcf_subset as ms
from farms_mujoco_b200 import models
from oracle.oracle import OraclePhysics


def _mjcf(body_xml, option='timestep="0.001" gravity="0 0 -9.81"', extra=''):
    return f'''<mujoco model="t">
  <compiler angle="radian" inertiafromgeom="false"/>
  <option {option} integrator="Euler" cone="pyramidal"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" friction="1 0 0" contype="1" conaffinity="1" condim="3"/>
    {body_xml}
  </worldbody>
  {extra}
</mujoco>'''


def test_free_fall_discrete_recurrence():
    """Semi-implicit Euler: v_n = -g n h, z_n = z0 - g h^2 n(n+1)/2 (exact)."""
    xml = _mjcf('''<body name="ball" pos="0 0 5"><freejoint name="root"/>
      <inertial pos="0 0 0" mass="2" diaginertia="0.1 0.1 0.1"/>
      <geom name="g" type="sphere" size="0.1" contype="1" conaffinity="0" friction="1 0 0"/></body>''')
    phys = OraclePhysics(ms.parse_mjcf(xml))
    n, h, g = 200, 1e-3, 9.81
    phys.step(n)
    assert phys.data.qvel[2] == pytest.approx(-g*n*h, rel=1e-12)
    assert phys.data.qpos[2] == pytest.approx(5 - g*h*h*n*(n + 1)/2, rel=1e-12)
    assert np.allclose(phys.data.qpos[3:7], [1, 0, 0, 0])


def test_hinge_pendulum_matches_closed_form_map():
    """Point-mass pendulum (fixed base): qacc = -(g/l) sin(q); q' integrates the new velocity."""
    xml = _mjcf('''<body name="base" pos="0 0 2"><body name="arm" pos="0 0 0">
      <joint name="j" type="hinge" axis="0 1 0" pos="0 0 0" damping="0" limited="false"/>
      <inertial pos="0 0 -0.5" mass="1" diaginertia="1e-9 1e-9 1e-9"/></body></body>''')
    model = ms.parse_mjcf(xml)
    phys = OraclePhysics(model)
    phys.data.qpos[0] = 0.3
    q, v, h = 0.3, 0.0, 1e-3
    for _ in range(300):
        phys.step()
        acc = -(9.81*0.5)/(0.25 + 1e-9)*np.sin(q)
        v += h*acc
        q += h*v
    assert phys.data.qpos[0] == pytest.approx(q, rel=1e-9)
    assert phys.data.qvel[0] == pytest.approx(v, rel=1e-9)


@pytest.mark.parametrize('name', ['swimmer8', 'salamander', 'centipede'])
def test_crb_mass_matrix_equals_jacobian_definition(name):
    """CRB (orc_crb) == sum_b m Jp'Jp + Jr' I Jr at a random configuration."""
    spec = models.MODELS[name]()
    model = ms.parse_mjcf(spec.mjcf)
    phys = OraclePhysics(model)
    rng = np.random.default_rng(1)
    phys.data.qpos[7:] = rng.uniform(-0.5, 0.5, model.nq - 7)
    quat = rng.normal(size=4)
    phys.data.qpos[3:7] = quat/np.linalg.norm(quat)
    phys.forward()
    dense, _ = ms.dense_mass_matrix(model, phys.data.qpos.copy())
    crb = phys.full_mass_matrix()
    assert np.abs(crb - dense).max() < 1e-12*np.abs(dense).max()
    # L'DL solve residual
    rhs = rng.normal(size=model.nv)
    assert np.allclose(crb @ phys.data.qacc, phys.arrays['qfrc_smooth'] + phys.arrays['qfrc_constraint'],
                       rtol=1e-8, atol=1e-9*np.abs(phys.arrays['qfrc_smooth']).max()) or phys.nefc
    x = np.linalg.solve(crb, rhs)
    assert np.allclose(crb @ x, rhs)


def test_momentum_drift_is_first_order_in_timestep():
    """Free-floating chain, internal actuation only, no gravity: the continuous system
    conserves linear momentum exactly, so the drift of semi-implicit Euler must shrink
    linearly with the timestep (it does only if CRB, RNE bias and actuation agree)."""
    drift = []
    for dt, n in ((1e-3, 50), (1e-4, 500)):
        spec = models.swimmer8()
        xml = spec.mjcf.replace('gravity="0.0 0.0 -9.81"', 'gravity="0.0 0.0 0.0"')
        xml = xml.replace('timestep="0.001"', f'timestep="{dt}"')
        model = ms.parse_mjcf(xml)
        assert model.timestep == dt and not model.gravity.any()
        phys = OraclePhysics(model)
        phys.data.ctrl[:] = np.random.default_rng(2).uniform(-0.5, 0.5, model.nu)
        phys.step_raw(n)
        phys.forward()
        lin = phys.arrays['body_linvel'].reshape(-1, 3)
        drift.append(np.linalg.norm((model.body_mass[:, None]*lin).sum(axis=0)))
    assert drift[0]/drift[1] == pytest.approx(10.0, rel=0.02)
    assert drift[1] < 2e-3


def test_sphere_rests_on_plane_with_weight_balanced():
    """Resting sphere: total contact normal force = m g, penetration > 0, friction <= mu N."""
    xml = _mjcf('''<body name="ball" pos="0 0 0.1"><freejoint name="root"/>
      <inertial pos="0 0 0" mass="1.5" diaginertia="0.01 0.01 0.01"/>
      <geom name="g" type="sphere" size="0.1" contype="1" conaffinity="0" friction="1 0 0"/></body>''')
    phys = OraclePhysics(ms.parse_mjcf(xml))
    for _ in range(3000):
        phys.step()
    assert phys.ncon == 1
    force = phys.contact_force(0)
    assert force[0] == pytest.approx(1.5*9.81, rel=1e-6)
    assert phys.data.contact[0].dist < 0
    assert np.hypot(force[1], force[2]) <= 1.0*force[0] + 1e-9
    assert abs(phys.data.qvel[2]) < 1e-8


def test_solver_kkt_conditions():
    """At the solver's answer: f >= 0, f = -D min(0, J a - aref), M a = qfrc_smooth + J' f."""
    spec = models.salamander()
    model = ms.parse_mjcf(spec.mjcf)
    phys = OraclePhysics(model)
    rng = np.random.default_rng(5)
    phys.data.qpos[2] -= 0.004
    phys.data.qvel[:] = rng.uniform(-0.5, 0.5, model.nv)
    phys.forward()
    assert phys.nefc > 0
    efc = phys.efc()
    resid = efc['J'] @ phys.data.qacc - efc['aref']
    assert (efc['force'] >= 0).all()
    assert np.allclose(efc['force'], -efc['D']*np.minimum(0, resid), rtol=1e-9, atol=1e-12)
    lhs = phys.full_mass_matrix() @ phys.data.qacc
    rhs = phys.arrays['qfrc_smooth'] + efc['J'].T @ efc['force']
    assert np.abs(lhs - rhs).max() < 1e-8*max(1.0, np.abs(rhs).max())


def test_joint_limit_force_pushes_back():
    spec = models.swimmer8()
    model = ms.parse_mjcf(spec.mjcf)
    phys = OraclePhysics(model)
    phys.data.qpos[7] = 1.05   # beyond the +1.0 rad limit
    phys.data.ctrl[model.actuator_id('actuator_position_joint_0')] = 1.05   # no actuator pull-back
    phys.forward()
    assert phys.nefc == 1
    assert phys.arrays['jnt_limit_force'][1] > 0
    assert phys.arrays['qfrc_constraint'][6] < 0


def test_terminal_velocity_under_quadratic_drag():
    """Sinking link: v_t = sqrt(m g_eff / (|c| viscosity)) with buoyancy off (drag.pyx:83-88)."""
    from farms_mujoco_b200.data import AnimatData
    from farms_mujoco_b200.simulation.physics import FarmsTables
    from oracle import farms_oracle as fo
    spec = models.swimmer8()
    spec.arena_options.water.buoyancy = False
    spec.arena_options.water.height = 100.0
    for link in spec.animat_options.morphology.links:
        link.drag_coefficients = [[-4.0, -4.0, -4.0], [-1e-3, -1e-3, -1e-3]]
    model = ms.parse_mjcf(spec.mjcf)
    phys = OraclePhysics(model)
    data = AnimatData.from_sensors_names(1e-3, 1, spec.links_names, spec.joints_names,
                                         spec.contacts_names, spec.xfrc_names)
    maps = fo.make_maps(model, data)
    tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options,
                         spec.arena_options, spec.simulation_options.units)
    handler = fo.SwimmingHandlerOracle(data, tables)
    units = spec.simulation_options.units
    for _ in range(1200):   # ~0.7 m of sinking, well above the floor at z = -2
        fo.physics2data(phys, 0, data, maps, units)
        handler.step(0)
        fo.apply_xfrc(phys, data, 0, maps['sensors'], units)
        phys.step()
    mass = model.body_mass[2]
    assert abs(phys.data.qvel[2]) == pytest.approx(np.sqrt(mass*9.81/4.0), rel=2e-3)


def test_buoyancy_matches_formula():
    """drag.pyx:139-149: lift = 1000 m g / rho_link * clamp((surface - z)/height, 0, 1)."""
    from oracle import farms_oracle as fo
    from farms_mujoco_b200.layout import sc
    links = np.zeros((1, 1, sc.link_size))
    links[0, 0, 2] = -0.004                      # 4 mm under the surface
    links[0, 0, [6, 13]] = 1.0                   # identity orientations (xyzw)
    xfrc = np.zeros((1, 1, 6))
    water = dict(surface=0.0, velocity=np.zeros(3), viscosity=1.0)
    done = fo.drag_forces(0, links, 0, xfrc, 0, np.zeros((2, 3)), water, mass=0.2, height=0.01,
                          density=800.0, gravity=-9.81, use_buoyancy=True)
    assert done
    assert xfrc[0, 0, 2] == pytest.approx(1000*0.2*9.81/800.0*0.4)
    links[0, 0, 2] = 0.5                          # above the surface: row untouched
    xfrc[:] = 7.0
    assert not fo.drag_forces(0, links, 0, xfrc, 0, np.zeros((2, 3)), water, mass=0.2, height=0.01,
                              density=800.0, gravity=-9.81, use_buoyancy=True)
    assert (xfrc == 7.0).all()


def test_plane_box_corners():
    """Plane-box collision (mjc_PlaneBox): a box foot flat on the floor touches with exactly its
    four lower corners, at the corners, with the corner height as distance; at rest the contact
    sensors carry the animat's weight."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from oracle.oracle import OraclePhysics
    spec = variant_models.salamander_box_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    qpos = np.array(model.key_qpos)
    qpos[2] -= 0.004                       # feet 2 mm into the floor
    orc.data.qpos[:] = qpos
    orc.forward()
    n = orc.ncon
    cand = orc.arrays['con_cand'][:n]
    ends = np.asarray(model.cand_end)[cand]
    geoms = np.asarray(model.cand_geom2)[cand]
    pos = orc.arrays['con_pos'].reshape(-1, 3)[:n]
    dist = orc.arrays['con_dist'][:n]
    feet = [g for g in range(model.ngeom) if model.geom_names[g].endswith('_foot')]
    assert len(feet) == 4
    for g in feet:
        sel = geoms == g
        assert sel.sum() == 4, (model.geom_names[g], sel.sum())
        assert np.all(ends[sel] >= 2) and np.all((ends[sel] - 2) & 4 == 0)      # the four z- corners
        assert np.allclose(dist[sel], dist[sel][0], atol=1e-12) and dist[sel][0] < 0
        assert np.allclose(pos[sel][:, 2], 0.5*dist[sel][0], atol=1e-12)        # halfway into the floor
        # the corners are those of the foot's 28 x 18 mm rectangle, turned by 0.3 rad about z
        span = pos[sel][:, :2].max(axis=0) - pos[sel][:, :2].min(axis=0)
        cs, sn = np.cos(0.3), np.sin(0.3)
        assert np.allclose(span, [0.028*cs + 0.018*sn, 0.028*sn + 0.018*cs], atol=1e-9)


def test_plane_ellipsoid_support_point():
    """Plane-ellipsoid collision (mjc_PlaneConvex with the ellipsoid's support function): the one
    contact of an ellipsoid foot is its lowest point -- on the surface, with the surface normal
    along the plane normal, and no sampled surface point lower -- and sits halfway between that
    point and the plane."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from farms_mujoco_b200.mjcf_subset import quat2mat
    spec = variant_models.salamander_ellipsoid_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    rng = np.random.default_rng(9)
    qpos = np.array(model.key_qpos)
    qpos[2] -= 0.012
    qpos[3:7] = [0.999, 0.02, -0.03, 0.03]              # the trunk tilted: no principal axis is vertical
    qpos[3:7] /= np.linalg.norm(qpos[3:7])
    qpos[7:] += rng.uniform(-0.1, 0.1, model.nq - 7)
    orc.data.qpos[:] = qpos
    orc.forward()
    n = orc.ncon
    cand = orc.arrays['con_cand'][:n]
    ends = np.asarray(model.cand_end)[cand]
    assert (ends == 10).sum() >= 2
    pos = orc.arrays['con_pos'].reshape(-1, 3)[:n]
    dist = orc.arrays['con_dist'][:n]
    gpos = orc.arrays['geom_xpos'].reshape(-1, 3)
    gmat = orc.arrays['geom_xmat'].reshape(-1, 3, 3)
    normal = np.array([0.0, 0.0, 1.0])
    u, v = np.meshgrid(np.linspace(0, np.pi, 400), np.linspace(0, 2*np.pi, 800))
    sphere = np.stack([np.sin(u)*np.cos(v), np.sin(u)*np.sin(v), np.cos(u)], axis=-1).reshape(-1, 3)
    for i in np.flatnonzero(ends == 10):
        g = int(np.asarray(model.cand_geom2)[cand[i]])
        size, R, centre = np.asarray(model.geom_size).reshape(-1, 3)[g], gmat[g], gpos[g]
        point = pos[i] + 0.5*dist[i]*normal              # the support point itself
        local = R.T @ (point - centre)
        assert abs(np.sum((local/size)**2) - 1.0) < 1e-12                      # on the surface
        grad = R @ (local/size**2)
        assert np.allclose(grad/np.linalg.norm(grad), -normal, atol=1e-12)     # lowest point: normal = -n
        assert abs(point @ normal - dist[i]) < 1e-12                           # plane z = 0
        heights = (centre + (sphere*size) @ R.T) @ normal
        assert heights.min() >= point @ normal - 1e-12


def test_plane_cylinder_points():
    """Plane-cylinder collision (mjc_PlaneCylinder): every contact point lies on a rim of the
    cylinder; the first is its lowest point (no sampled surface point is lower), the second the rim
    point straight along the axis from it, the last two sit on the lower rim 120 degrees either side
    of the first; each contact is halfway between its point and the plane."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    spec = variant_models.salamander_cylinder_feet()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    rng = np.random.default_rng(9)
    qpos = np.array(model.key_qpos)
    qpos[2] -= 0.02
    qpos[3:7] = [0.999, 0.02, -0.03, 0.03]
    qpos[3:7] /= np.linalg.norm(qpos[3:7])
    qpos[7:] += rng.uniform(-0.1, 0.1, model.nq - 7)
    orc.data.qpos[:] = qpos
    orc.forward()
    n = orc.ncon
    cand = orc.arrays['con_cand'][:n]
    ends = np.asarray(model.cand_end)[cand]
    geoms = np.asarray(model.cand_geom2)[cand]
    pos = orc.arrays['con_pos'].reshape(-1, 3)[:n]
    dist = orc.arrays['con_dist'][:n]
    gpos = orc.arrays['geom_xpos'].reshape(-1, 3)
    gmat = orc.arrays['geom_xmat'].reshape(-1, 3, 3)
    normal = np.array([0.0, 0.0, 1.0])
    checked = 0
    for g in sorted(set(geoms[ends >= 11].tolist())):
        radius, half = np.asarray(model.geom_size).reshape(-1, 3)[g][:2]
        R, centre = gmat[g], gpos[g]
        points = {}
        for i in np.flatnonzero((geoms == g) & (ends >= 11)):
            point = pos[i] + 0.5*dist[i]*normal
            assert abs(point @ normal - dist[i]) < 1e-12
            local = R.T @ (point - centre)
            assert abs(np.hypot(local[0], local[1]) - radius) < 1e-12 and abs(abs(local[2]) - half) < 1e-12    # on a rim
            points[int(ends[i]) - 11] = local
        assert 0 in points                                   # the lowest point is always among them
        u, v = np.meshgrid(np.linspace(0, 2*np.pi, 2000), [-half, half])
        rim = np.stack([radius*np.cos(u), radius*np.sin(u), v], axis=-1).reshape(-1, 3)
        assert ((centre + rim @ R.T) @ normal).min() >= (centre + R @ points[0]) @ normal - 1e-9
        if 1 in points:
            assert np.allclose(points[1][:2], points[0][:2], atol=1e-12) and np.isclose(points[1][2], -points[0][2])
        for k in (2, 3):
            if k in points:
                assert np.isclose(points[k][2], points[0][2])
                cosang = points[k][:2] @ points[0][:2]/radius**2
                assert np.isclose(cosang, -0.5, atol=1e-12)
                checked += 1
    assert checked >= 2


def test_pair_sphere_sphere_contact():
    """Explicit <pair>, sphere-sphere (mjc_SphereSphere): the contact sits half way between the two
    surfaces on the line of centres, the normal runs from geom1 to geom2, the distance is the gap
    between the surfaces; and the contact is an INTERNAL force -- the rows' Jacobian is the
    difference of the two bodies' point Jacobians (mj_jacDifPair), so the constraint force has no
    component on the six dofs of the floating root (Newton's third law)."""
    import variant_models
    from farms_mujoco_b200 import mjcf_subset
    from oracle.oracle import OraclePhysics
    spec = variant_models.salamander_foot_pairs()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    orc = OraclePhysics(model)
    orc.reset(keyframe_id=0)
    orc.data.qpos[:] = variant_models.folded_legs_qpos(model, 1.0)
    orc.forward()
    assert orc.ncon == 2
    cand = orc.arrays['con_cand'][:2]
    assert (np.asarray(model.cand_end)[cand] == 20).all()
    radius = 0.017
    for i, c in enumerate(cand):
        g1, g2 = model.cand_geom1[c], model.cand_geom2[c]
        gx = orc.arrays['geom_xpos'].reshape(-1, 3)
        p1, p2 = gx[g1], gx[g2]
        gap = np.linalg.norm(p2 - p1) - 2*radius
        normal = (p2 - p1)/np.linalg.norm(p2 - p1)
        assert gap < 0 and abs(orc.arrays['con_dist'][i] - gap) < 1e-14
        assert np.allclose(orc.arrays['con_frame'].reshape(-1, 9)[i, :3], normal, atol=1e-14)
        assert np.allclose(orc.arrays['con_pos'].reshape(-1, 3)[i], 0.5*(p1 + p2), atol=1e-14)   # equal radii
        force = orc.contact_force(i)
        assert force[0] > 0.1                                        # the feet push apart
    qfrc = orc.arrays['qfrc_constraint']
    assert np.abs(qfrc[6:]).max() > 1e-3
    assert np.abs(qfrc[:6]).max() < 1e-9*np.abs(qfrc[6:]).max()


def test_instrumented_op_count_matches_stock_oracle(tmp_path):
    """oracle/opcount: the oracle compiled with the counting scalar reproduces the stock oracle's
    bits and reports the committed per-step operation counts (profiles/oracle_opcount.json,
    BASELINE.md section 3) for the contact-free models, whose count does not depend on the state."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'oracle', 'opcount', 'count_ops.py'),
                          '--steps', '6', '--warm', '2', '--envs', '1'], check=True, capture_output=True, text=True)
    got = json.loads(out.stdout)['models']
    with open(os.path.join(root, 'profiles', 'oracle_opcount.json')) as f:
        committed = json.load(f)['models']
    for name, rec in got.items():
        assert rec['same_bits_as_stock_oracle'], name
    for name in ('swimmer8', 'salamander_swim'):
        for key in ('add', 'mul', 'div', 'sqrt', 'flop'):
            assert got[name][key] == committed[name][key], (name, key, got[name][key], committed[name][key])
    assert got['centipede']['flop'] > got['salamander']['flop'] > got['swimmer8']['flop']
