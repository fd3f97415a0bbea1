"""Instrumented op-count of the fp64 oracle (TEST INFRASTRUCTURE; SURVEY.md section 8(d):
"define algorithmic FLOPs per env-step by an instrumented oracle count").

Compiles oracle/mjstep_oracle.c as C++ with every `double` of the oracle's own code replaced
by the counting scalar of counted.h (same layout, so oracle.py drives it unchanged), rolls
each BASELINE model for a few hundred steps from the bench's synthetic initial states and
prints / writes the mean operation counts of one mj_step (orc_step): the restated MuJoCo
pipeline -- kinematics, CRB, L'DL, collision, constraint rows, RNE, Newton solver, Euler.  The
NumPy glue (physics2data, drag: ~250 flop per swimming link) is listed apart.

    python oracle/opcount/count_ops.py [--steps 200] [--envs 4] > oracle/opcount/opcount.json
"""
import argparse
import ctypes as ct
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
NAMES = ['add', 'mul', 'div', 'sqrt', 'transcendental', 'compare']


def build():
    src = open(os.path.join(HERE, '..', 'mjstep_oracle.c')).read()
    hdr = open(os.path.join(HERE, '..', 'oracle.h')).read()
    # the oracle's own doubles count; the ABI header (FbModel) keeps its types
    src = src.replace('#include "oracle.h"', hdr)
    head, body = src.split('#include "../include/farms_b200.h"', 1)
    # FbModel's double arrays become cnt_t arrays too (same layout): model constants enter the
    # arithmetic as counted operands
    abi = open(os.path.join(ROOT, 'include', 'farms_b200.h')).read()
    body = re.sub(r'\bdouble\b', 'cnt_t', abi + body)
    gen = os.path.join(HERE, '_counted_oracle.cpp')
    with open(gen, 'w') as f:
        f.write('#include "counted.h"\nlong long g_opc[OPC_N];\nextern "C" {\n' + body +
                '\nvoid orc_opcount(long long *out, int reset) { for (int i = 0; i < OPC_N; i++) '
                '{ out[i] = g_opc[i]; if (reset) g_opc[i] = 0; } }\n}\n')
    so = os.path.join(HERE, '_libcounted.so')
    subprocess.run(['g++', '-O1', '-fPIC', '-shared', '-fpermissive', '-w', '-std=c++17', '-I', HERE,
                    '-o', so, gen], check=True)
    return so


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--envs', type=int, default=4)
    ap.add_argument('--warm', type=int, default=8, help='uncounted steps after the reset')
    args = ap.parse_args()
    from oracle import oracle as orc
    counted = ct.CDLL(build())
    plain = orc.lib()                       # sets argtypes on the stock library ...
    for name in ('orc_kinematics', 'orc_com_pos', 'orc_crb', 'orc_collision', 'orc_make_constraint',
                 'orc_com_vel', 'orc_passive', 'orc_actuation', 'orc_solve_constraints', 'orc_forward',
                 'orc_euler', 'orc_step'):
        getattr(counted, name).argtypes = [ct.c_void_p, ct.c_void_p]
        getattr(counted, name).restype = None
    counted.orc_step_n.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int]
    counted.orc_step_n.restype = None
    counted.orc_opcount.argtypes = [ct.POINTER(ct.c_longlong), ct.c_int]
    orc._LIB = counted                      # ... and the counting one takes its place  # pylint: disable=protected-access
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.data import AnimatData
    from farms_mujoco_b200.simulation.physics import FarmsTables
    from farms_mujoco_b200.sharding import synthetic_inputs
    from farms_mujoco_b200.models import travelling_wave_parameters
    from oracle import farms_oracle as fo
    out = {}
    buf = (ct.c_longlong*6)()
    for name in ('swimmer8', 'salamander_swim', 'salamander', 'centipede'):
        spec = models.MODELS[name]()
        model = mjcf_subset.parse_mjcf(spec.mjcf)
        qpos0, _, phase = synthetic_inputs(model, np.arange(args.envs))
        joints, amp, freq, lag = travelling_wave_parameters(spec)
        acts = np.array([model.actuator_id(f'actuator_position_{j}') for j in joints])
        total, niter, ncon, nefc = np.zeros(6), 0, 0, 0
        final = {}
        for e in list(range(args.envs)) + [-1]:
            if e < 0:
                # environment 0 once more on the stock (uncounted) library: same bits expected
                orc._LIB, e = plain, 0  # pylint: disable=protected-access
                counted.orc_opcount(buf, 1)
            # the loop of bench.py's CPU arm: sensors -> swimming -> control -> mj_step
            physics = orc.OraclePhysics(model)
            data = AnimatData.from_sensors_names(model.timestep, 64, spec.links_names, spec.joints_names,
                                                 spec.contacts_names, spec.xfrc_names)
            maps = fo.make_maps(model, data)
            tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options,
                                 spec.arena_options, spec.simulation_options.units)
            handler = fo.SwimmingHandlerOracle(data, tables)
            units = spec.simulation_options.units
            physics.reset(keyframe_id=0)
            physics.data.qpos[:] = qpos0[e]
            physics.forward()
            for k in range(args.warm + args.steps):
                if k == args.warm:
                    counted.orc_opcount(buf, 1)
                row = k % 64
                data.sensors.contacts.array[row] = 0
                data.sensors.joints.array[row] = 0
                fo.physics2data(physics, row, data, maps, units)
                if len(tables.swim_links_index):
                    handler.step(row)
                    fo.apply_xfrc(physics, data, row, maps['sensors'], units)
                physics.data.ctrl[acts] = amp*np.sin(2*np.pi*freq*k*model.timestep - lag + phase[e])
                physics.step()
                if k >= args.warm and orc._LIB is counted:  # pylint: disable=protected-access
                    niter += physics.solver_niter
                    ncon += physics.ncon
                    nefc += physics.nefc
            counted.orc_opcount(buf, 1)
            total += np.array(list(buf), dtype=float)
            final.setdefault(e, []).append(np.concatenate([physics.data.qpos, physics.data.qvel]).copy())
        orc._LIB = counted  # pylint: disable=protected-access
        n = args.envs*args.steps
        per = total/n
        rec = {k_: float(round(v, 1)) for k_, v in zip(NAMES, per)}
        rec['flop'] = float(round(per[:5].sum(), 1))
        rec.update(nv=int(model.nv), nbody=int(model.nbody), mean_contacts=round(ncon/n, 2),
                   mean_constraint_rows=round(nefc/n, 2), mean_newton_iterations=round(niter/n, 2),
                   swimming_links=int(len(tables.swim_links_index)),
                   same_bits_as_stock_oracle=bool(np.array_equal(final[0][0], final[0][1])))
        out[name] = rec
    json.dump({'unit': 'fp64 operations per mj_step of one environment (mean over %d envs x %d steps)' % (args.envs, args.steps),
               'flop_definition': 'add + mul + div + sqrt + transcendental, one each (no FMA contraction); compares listed apart',
               'models': out}, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main()
