#!/usr/bin/env python
"""Generate tests/golden/*.npz from the fp64 CPU oracle.

The reference itself cannot run here (MuJoCo, dm_control and farms_core are
absent), so these vectors pin the ORACLE: they freeze its
outputs at the commit that generated them, guard against silent drift of the
restatement, and give the GPU tests fixed inputs/outputs that do not need the
oracle library.  Regenerate only on a deliberate oracle change:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import make_case, oracle_rollout  # noqa: E402
from farms_mujoco_b200.data import AnimatData  # noqa: E402
from farms_mujoco_b200.simulation.physics import FarmsTables  # noqa: E402
from oracle import farms_oracle as fo  # noqa: E402

CASES = [('swimmer8', 16), ('salamander_swim', 12), ('salamander', 12), ('centipede', 8)]


def main():
    for name, n_rows in CASES:
        spec, model, qpos0, qvel0, ctrl = make_case(name, 2, seed=7)
        data = AnimatData.from_sensors_names(model.timestep, 1, spec.links_names, spec.joints_names,
                                             spec.contacts_names, spec.xfrc_names)
        maps = fo.make_maps(model, data)
        tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options,
                             spec.arena_options, spec.simulation_options.units)
        out = dict(qpos0=qpos0, qvel0=qvel0, ctrl=ctrl, n_rows=n_rows)
        for env in range(2):
            _, log, states = oracle_rollout(spec, model, tables, n_rows, qpos0[env], qvel0[env], ctrl[env])
            out[f'qpos_{env}'] = states[-1][0]
            out[f'qvel_{env}'] = states[-1][1]
            for kind in ('links', 'joints', 'contacts', 'xfrc'):
                out[f'{kind}_{env}'] = getattr(log.sensors, kind).array
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
