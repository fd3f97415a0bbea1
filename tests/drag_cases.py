"""Seeded link rows for the stand-alone drag operator (fb_drag_forces) and the oracle's answer
(oracle/farms_oracle.py drag_forces, the restatement of drag.pyx:152-268)."""

import numpy as np

from oracle import farms_oracle as fo


def make_rows(n, seed=7, surface=0.05):
    """n link rows around the water surface: random unit quaternions (CoM and URDF frames differ),
    velocities of both signs including exact zeros, masses including 0, a few rows exactly at the
    surface and above it."""
    rng = np.random.default_rng(seed)
    links = np.zeros((n, 20))
    links[:, 0:3] = rng.uniform(-1, 1, (n, 3))*[1.0, 1.0, 0.2]
    for c in (3, 10):
        q = rng.normal(size=(n, 4))
        links[:, c:c+4] = q/np.linalg.norm(q, axis=1, keepdims=True)
    links[:, 7:10] = links[:, 0:3] + rng.uniform(-0.01, 0.01, (n, 3))
    links[:, 14:20] = rng.uniform(-2, 2, (n, 6))
    links[::7, 14] = 0.0
    links[::11, 19] = 0.0
    if n > 3:
        links[1, 2] = surface            # exactly at the surface: applied, no buoyancy
        links[2, 2] = surface + 1e-3     # above: untouched
        links[3, 2] = surface - 1e-4     # shallow: partial buoyancy
    coef = -rng.uniform(0.0, 2.0, (n, 2, 3))
    mass = rng.uniform(0.0, 0.5, n)
    mass[::5] = 0.0
    height = rng.uniform(0.005, 0.05, n)
    density = rng.uniform(500.0, 1500.0, n)
    return dict(links=links, coef=coef, mass=mass, height=height, density=density, surface=surface,
                wvel=np.array([0.3, -0.2, 0.1]), viscosity=1.3, gravity=-9.81)


def oracle_answer(case, use_buoyancy, xfrc0):
    n = case['links'].shape[0]
    xfrc = xfrc0.copy()[None]
    links = case['links'][None]
    water = {'surface': case['surface'], 'velocity': case['wvel'], 'viscosity': case['viscosity']}
    applied = np.zeros(n, dtype=bool)
    for i in range(n):
        applied[i] = fo.drag_forces(0, links, i, xfrc, i, case['coef'][i], water, case['mass'][i],
                                    case['height'][i], case['density'][i], case['gravity'], use_buoyancy)
    return xfrc[0], applied


def check_operator(library, n=257):
    """fb_drag_forces through the C ABI against the oracle, with and without buoyancy; rows above
    the surface keep the values the caller passed in."""
    from farms_mujoco_b200.engine import drag_forces_rows
    case = make_rows(n)
    worst = 0.0
    for use_buoyancy in (True, False):
        xfrc = np.random.default_rng(3).uniform(-1, 1, (n, 6))
        ref, ref_applied = oracle_answer(case, use_buoyancy, xfrc)
        applied = drag_forces_rows(case['links'], case['coef'], case['mass'], case['height'], case['density'],
                                   case['surface'], case['wvel'], case['viscosity'], case['gravity'],
                                   use_buoyancy, xfrc, library=library)
        assert (applied == ref_applied).all()
        assert not applied[2] and applied[1] and applied[3]
        assert (xfrc[~applied] == ref[~applied]).all()          # untouched, bit for bit
        scale = np.maximum(1.0, np.abs(ref))
        worst = max(worst, float((np.abs(xfrc - ref)/scale).max()))
    return worst
