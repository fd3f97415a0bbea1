#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s of the fused FARMS step on B200.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W            our arm (CUDA)
  python bench.py --impl reference --gpus N --steps K ...   CPU arm (oracle port)

Workload (config.workload): swimming salamander (28 links, nv=33) with
hydrodynamic drag + buoyancy and full sensor logging (links/joints/contacts/
xfrc), the 65,536-environment sweep BASELINE.json's metric is quoted on
(configs[4]): 65,536 environments per GPU, N GPUs step N x 65,536 independent
environments (weak scaling, no data-path collective).  ``--envs-per-gpu 16384`` is
configs[2].  One bench "step" = one launch of the fused kernels advancing every
environment by ``--inner`` physics steps.

value  : device-timed (CUDA events on the engine's stream, max over ranks),
         state resident in HBM.
e2e    : the same through BatchedPhysics.step_host on pinned HOST buffers:
         ctrl uploaded, last links+joints log row of every env downloaded.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

_JSON_OUT = sys.stdout

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec (box, device-timed) at 1/2/4/8 B200 vs ref CPU on host cores'
WORKLOAD = 'salamander_swim'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs-per-gpu', type=int, default=65536,
                    help='BASELINE.json config 5 at N = 1: 65,536 swimming salamanders on one B200 (config 3 is '
                         '--envs-per-gpu 16384)')
    ap.add_argument('--inner', type=int, default=16, help='physics steps per launch')
    ap.add_argument('--ring', type=int, default=64, help='log ring rows per env')
    ap.add_argument('--model', default=WORKLOAD)
    ap.add_argument('--team', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-fast', action='store_true', help='team kernel only (comparison runs)')
    ap.add_argument('--team-constraints', action='store_true',
                    help='hand-overs (limits / contacts) go to the team kernel instead of the per-thread '
                         'constrained kernel (comparison runs)')
    return ap.parse_args()


# ------------------------------------------------------------------ inputs
from farms_mujoco_b200.sharding import synthetic_inputs, gather_env_statistics, bind_host_to_device  # noqa: E402


def wave_controller(spec, model):
    from farms_mujoco_b200.models import travelling_wave_parameters
    joints, amp, freq, lag = travelling_wave_parameters(spec)
    acts = [model.actuator_id(f'actuator_position_{j}') for j in joints]
    return acts, amp, freq, lag


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:  # pylint: disable=broad-except
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:  # pylint: disable=broad-except
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:  # pylint: disable=broad-except
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {
            'sm_mhz': float(np.median(self.samples)) if self.samples else None,
            'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons),
        }


# ---------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process = one environment of the workload on the oracle port, full
    reference order per step (sensors -> swimming -> control -> mj_step)."""
    model_name, seconds, seed = args
    sys.path.insert(0, ROOT)
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.data import AnimatData
    from farms_mujoco_b200.simulation.physics import FarmsTables
    from oracle.oracle import OraclePhysics
    from oracle import farms_oracle as fo
    spec = models.MODELS[model_name]()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    physics = OraclePhysics(model)
    n_rows = 256
    data = AnimatData.from_sensors_names(model.timestep, n_rows, spec.links_names,
                                         spec.joints_names, spec.contacts_names, spec.xfrc_names)
    maps = fo.make_maps(model, data)
    tables = FarmsTables(model, data.sensors, maps['sensors'], spec.animat_options,
                         spec.arena_options, spec.simulation_options.units)
    handler = fo.SwimmingHandlerOracle(data, tables)
    units = spec.simulation_options.units
    qpos, qvel, phase = synthetic_inputs(model, np.array([seed]))
    joints, amp, freq, lag = wave_controller(spec, model)
    acts = np.array(joints)
    physics.reset(keyframe_id=0)
    physics.data.qpos[:] = qpos[0]
    physics.forward()
    steps, t0 = 0, time.perf_counter()
    while True:
        row = steps % n_rows
        data.sensors.contacts.array[row] = 0
        data.sensors.joints.array[row] = 0
        fo.physics2data(physics, row, data, maps, units)
        if len(tables.swim_links_index):
            handler.step(row)
            fo.apply_xfrc(physics, data, row, maps['sensors'], units)
        t = steps*model.timestep
        physics.data.ctrl[acts] = amp*np.sin(2*np.pi*freq*t - lag + phase[0])
        physics.step()
        steps += 1
        if steps % 16 == 0 and time.perf_counter() - t0 > seconds:
            break
    return steps, time.perf_counter() - t0


def cpu_baseline(model_name, seconds, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(model_name, seconds, i) for i in range(cores)])
    rate = sum(s/t for s, t in res)
    return {
        'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
        'sample': (f'{cores} processes x 1 env of {model_name}, ~{seconds:.0f} s each, fp64 C oracle '
                   'step + NumPy physics2data/contacts/drag port (CPU restatement, not MuJoCo)'),
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    base = cpu_baseline(args.model, max(2.0, args.cpu_seconds))
    n_envs = args.envs_per_gpu*args.gpus
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'env-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3*n_envs*args.inner/base['value'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.model} x {n_envs} envs (bounded sample: {base["sample"]})',
                   'physics_steps_per_step': args.inner},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), file=_JSON_OUT, flush=True)


# ---------------------------------------------------------------- GPU arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.engine import BatchedPhysics
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the engine has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    # host side of the e2e arm on the GPU's own NUMA node (pinned buffers, controller threads)
    bound = None if os.environ.get('FARMS_B200_NO_BIND') else bind_host_to_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    spec = models.MODELS[args.model]()
    model = mjcf_subset.parse_mjcf(spec.mjcf)
    n_local = args.envs_per_gpu
    env_ids = np.arange(rank*n_local, (rank + 1)*n_local)
    qpos0, qvel0, phase = synthetic_inputs(model, env_ids)
    physics = BatchedPhysics.from_spec(spec, n_local, buffer_size=args.ring, device=local_rank,
                                       team_lanes=args.team)
    physics.set_env_phase(phase)
    physics.set_wave_controller(*wave_controller(spec, model))
    if args.no_fast:
        physics.set_fast_path(False)
    if args.team_constraints:
        physics.set_constraint_path(False)
    physics.reset(qpos0, qvel0)

    stream_ptr = __import__('ctypes').c_void_p()
    physics._check(physics.lib.fb_device_ptr_stream(physics._handle, __import__('ctypes').byref(stream_ptr)))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            ev0.record()
            for _ in range(steps):
                fn()
            ev1.record()
        physics.synchronize()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm
    for _ in range(args.warmup):
        physics.step(args.inner, sync=False)
    physics.synchronize()
    launches0 = physics.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = []

    def one():
        physics.step(args.inner, sync=False)

    total_ms = timed(one, args.steps)
    # per-launch kernel time from the engine's own events (last launch)
    kernel_ms.append(physics.last_step_ms())
    clocks = sampler.summary()
    env_steps = n_local*world*args.inner*args.steps
    value = env_steps/(total_ms*1e-3)
    launches = physics.launch_count() - launches0
    handed_over = physics.last_pending if physics.fast_path else n_local

    # ---- end-to-end arm on pinned host buffers
    e2e = None
    if not args.no_e2e:
        nl, njf = len(spec.links_names), len(spec.joints_names)
        # three sets of pinned host buffers: the engine's pipelined step_host overlaps the
        # download of launch i with the kernels of launch i+1 (include/farms_b200.h); the third
        # set lets the host enqueue launch i+1 (and its ctrl upload) while launch i still runs
        NSETS = 3
        ctrl_host = [torch.zeros((n_local, model.nu), dtype=torch.float32).pin_memory() for _ in range(NSETS)]
        links_host = [torch.empty((n_local, nl, 20), dtype=torch.float32).pin_memory() for _ in range(NSETS)]
        # the joints row comes down as the four columns the path writes (position, velocity,
        # torque, limit force); physics.py:481-524 leaves the other 14 of the 18 zero
        from farms_mujoco_b200.layout import sc
        jcols = [sc.joint_position, sc.joint_velocity, sc.joint_torque, sc.joint_limit_force]
        physics.set_host_joint_columns(jcols)
        joints_host = [torch.empty((n_local, njf, len(jcols)), dtype=torch.float32).pin_memory() for _ in range(NSETS)]
        physics.set_wave_controller(None, None, None, None)   # ctrl now comes from the host
        acts, amp, freq, lag = wave_controller(spec, model)
        acts_t = torch.as_tensor(np.array(acts))
        amp_t, freq_t = torch.as_tensor(amp, dtype=torch.float32), torch.as_tensor(freq, dtype=torch.float32)
        lag_t = torch.as_tensor(lag, dtype=torch.float32)
        phase_t = torch.as_tensor(phase, dtype=torch.float32)[:, None]
        # amp sin(a - lag + phase) = [amp sin(a - lag)] cos(phase) + [amp cos(a - lag)] sin(phase):
        # the per-environment factors are constant, so a call costs two multiply-adds per entry
        cos_ph, sin_ph = torch.cos(phase_t), torch.sin(phase_t)
        wave_buf = torch.empty((n_local, len(acts)), dtype=torch.float32)
        # where the wave goes in ctrl: a strided view when the actuator ids are evenly spaced (they
        # are: position / velocity / motor per joint), a column scatter otherwise
        acts_np = np.asarray(acts)
        stride = int(acts_np[1] - acts_np[0]) if len(acts_np) > 1 else 1
        regular = len(acts_np) > 1 and stride > 0 and bool(np.all(np.diff(acts_np) == stride))
        ctrl_view = [torch.as_strided(c, (n_local, len(acts)), (model.nu, stride), int(acts_np[0])) if regular else None
                     for c in ctrl_host]
        if world > 1:
            # torchrun pins every rank to one OpenMP thread; the host-side controller may use its share
            torch.set_num_threads(max(1, min(len(bound) if bound else 1 << 30, (os.cpu_count() or world)//world)))
        checksum = [0.0]
        calls = [0]
        pending = [None]*NSETS

        def one_host():
            # host-side controller of this outer iteration (task.py:288-321 analogue):
            # the same travelling wave, evaluated on the host and uploaded
            k = calls[0] % NSETS
            calls[0] += 1
            if pending[k] is not None:
                # the buffers of NSETS calls ago are complete once this call is allowed to reuse
                # them: consume the result (every step's rows are read on the host)
                physics.host_wait_call(pending[k])
                checksum[0] += float(links_host[k][0, 0, 0]) + float(joints_host[k][-1, 0, 0])
            t = physics.iteration*model.timestep
            arg = 2*np.pi*freq_t*t - lag_t
            torch.mul(cos_ph, amp_t*torch.sin(arg), out=wave_buf)
            wave_buf.addcmul_(sin_ph, amp_t*torch.cos(arg))
            if regular:
                ctrl_view[k].copy_(wave_buf)
            else:
                ctrl_host[k].index_copy_(1, acts_t, wave_buf)
            pending[k] = physics.step_host(args.inner, ctrl=ctrl_host[k], links_row=links_host[k],
                                           joints_row=joints_host[k], pipelined=True)

        for _ in range(max(NSETS, args.warmup)):
            one_host()
        physics.host_wait()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_host()
        physics.host_wait()
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], device='cuda')
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        links_host = links_host[(calls[0] - 1) % NSETS]
        ctrl_host, joints_host = ctrl_host[0], joints_host[0]
        e2e = {
            'value': env_steps/float(wall.item()), 'unit': 'env-steps/s',
            'h2d_bytes_per_step': int(ctrl_host.numel()*4*world),
            'd2h_bytes_per_step': int((links_host.numel() + joints_host.numel())*4*world),
            'checksum': float(links_host[:, 0, 0].double().sum()) + checksum[0],
            'pipelined': 'device->host copy of launch i overlaps the kernels of launch i+1 (fb_step_host_async, three host buffer sets); ctrl is fetched from pinned host memory by the SMs on an upload stream',
            'host_cpus_bound': len(bound) if bound else None,
            'rows_down': 'last links row [n_envs, n_links, 20] + joints row [n_envs, n_joints, 4 written columns]',
        }

    # ---- optional end-of-rollout gather of per-env statistics (NCCL)
    stats = torch.as_tensor(physics.qpos[:, :3].astype(np.float32), device='cuda')
    stats = gather_env_statistics(stats, world)
    all_flags = physics.flags
    flags = int(np.count_nonzero(all_flags & 1))
    solver_flags = int(np.count_nonzero(all_flags & 4))

    if rank == 0:
        b_log = physics.log_bytes_per_env_step
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        k_ms = kernel_ms[-1]
        achieved = n_local*args.inner*b_log/(k_ms*1e-3)/1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            # measured DRAM bytes per launch exist for the profiled (workload, batch, steps) only
            traffic = tj.get('by_workload', {}).get(f'{args.model}:{n_local}:{args.inner}', {}).get(
                'dram_bytes_per_launch')
        # FP32 side of the roofline ("report both"): arithmetic of the step against the
        # FFMA throughput measured on this device by the library's own micro-benchmark.  Two
        # numerators: the oracle's instrumented op count of one mj_step (profiles/oracle_opcount.json,
        # MuJoCo's CRB + L'DL pipeline in fp64 -- the shared "algorithmic" figure) and the FADD / FMUL /
        # FFMA the kernel executes (ncu, profiles/; the articulated-body recursion needs fewer).
        fp32 = None
        try:
            from farms_mujoco_b200.engine import measure_fp32_peak
            fp32_peak = measure_fp32_peak(local_rank)
            with open(os.path.join(ROOT, 'profiles', 'oracle_opcount.json')) as f:
                oc = json.load(f)
            rec = oc['models'].get(args.model)
            if rec:
                per_gpu = value/world
                fp32 = {
                    'peak_tflops': fp32_peak, 'peak_source': 'fb_measure_fp32_peak (FFMA micro-benchmark, this run)',
                    'oracle_flop_per_env_step': rec['flop'],
                    'oracle_equivalent_tflops': per_gpu*rec['flop']*1e-12,
                    'oracle_equivalent_frac': per_gpu*rec['flop']*1e-12/fp32_peak,
                    'executed_flop_per_env_step': oc.get('executed_flop_per_env_step', {}).get(args.model),
                }
                if fp32['executed_flop_per_env_step']:
                    fp32['executed_tflops'] = per_gpu*fp32['executed_flop_per_env_step']*1e-12
                    fp32['executed_frac'] = fp32['executed_tflops']/fp32_peak
        except (OSError, KeyError, ValueError) as exc:
            fp32 = {'error': str(exc)}
        out = {
            'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms/args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {
                'workload': (f'{args.model}: {n_local} envs/GPU x {world} GPU, drag+buoyancy, full '
                             'links/joints/contacts/xfrc log, on-device travelling-wave control'),
                'n_envs': n_local*world, 'physics_steps_per_step': args.inner, 'timestep': model.timestep,
                'nv': model.nv, 'nbody': model.nbody,
                'kernels': (('fb_fast_kernel (1 thread = 1 env, articulated-body recursion) then '
                             + ('fb_fastc_kernel (1 thread = 1 env, matrix-free Newton on the same recursion) '
                                'on the environments with an active limit / contact' if physics.constraint_path
                                else 'fb_step_kernel (lane team = 1 env, constraint solver) on the hand-overs'))
                            if physics.fast_path else 'fb_step_kernel only'),
                'fast_envs_per_block': physics.fast_path,
                'fast_slim_layout': bool(physics.fast_slim),
                'fast_warps_per_block': max(1, int(physics.fast_slim)),
                'fast_smem_bytes_per_env': physics.fast_smem_bytes_per_env,
                'handed_over_envs_last_launch': handed_over,
                'team_lanes': physics.team_lanes, 'team_smem_bytes_per_env': physics.smem_bytes_per_env,
                'l2_policy': (f'no flush: each step appends {n_local*args.inner*b_log/1e6:.0f} MB of new '
                              f'log rows to a {n_local*args.ring*b_log/1e9:.1f} GB ring (> 126 MB L2)'),
                'diverged_envs': flags, 'solver_pivot_clamped_envs': solver_flags,
            },
            'clocks': clocks,
            'e2e': e2e,
            'gpu_launches': int(launches),
            'fp32': fp32,
            'roofline': {
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved/peak, 'traffic': traffic,
                'kernel': ((f'fb_fastc_kernel<{physics.fast_path}>' if physics.constraint_path and 2*handed_over > n_local
                            else f'fb_fast_kernel<{physics.fast_path},{int(physics.fast_slim > 0)},{int(physics.fast_slim > 1)}>') if physics.fast_path and
                           (physics.constraint_path or 2*handed_over <= n_local)
                           else f'fb_step_kernel<{physics.team_lanes}>'),
                'kernel_ms': k_ms,
                'algorithmic_bytes_per_env_step': b_log,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650',
                'note': ('log-write bytes only (the one HBM stream of the step); kernel_ms is the CUDA-event '
                         'time of one fb_step launch pair on the engine stream; the step is FP32-issue/'
                         'latency bound, see profiles/ for issue-slot utilisation'),
            },
        }
        if not args.no_cpu_baseline and world == 1:
            out['cpu_baseline'] = cpu_baseline(args.model, args.cpu_seconds)
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner) write
    # to fd 1 too: everything but our line goes to stderr.
    global _JSON_OUT  # pylint: disable=global-statement
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit('launch N>1 with torch.distributed.run (one rank per GPU)')
    run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
