"""Where does the end-to-end arm's time go?  (GPU box; prints a few ms-per-launch figures)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from farms_mujoco_b200 import models, mjcf_subset
from farms_mujoco_b200.engine import BatchedPhysics
from farms_mujoco_b200.sharding import synthetic_inputs
from farms_mujoco_b200.layout import sc
import os
n, inner, reps = 65536, 16, (6 if os.environ.get('FARMS_B200_TRACE') else 20)
spec = models.MODELS['salamander_swim'](); model = mjcf_subset.parse_mjcf(spec.mjcf)
qpos0, qvel0, phase = synthetic_inputs(model, np.arange(n))
ph = BatchedPhysics.from_spec(spec, n, buffer_size=64)
ph.set_env_phase(phase); ph.reset(qpos0, qvel0)
nl, nj = len(spec.links_names), len(spec.joints_names)
jc = [sc.joint_position, sc.joint_velocity, sc.joint_torque, sc.joint_limit_force]
ph.set_host_joint_columns(jc)
NS = int(os.environ.get('NSETS', 3))
ctrl = [torch.zeros((n, model.nu), dtype=torch.float32).pin_memory() for _ in range(NS)]
links = [torch.empty((n, nl, 20), dtype=torch.float32).pin_memory() for _ in range(NS)]
joints = [torch.empty((n, nj, 4), dtype=torch.float32).pin_memory() for _ in range(NS)]
# raw PCIe figures
dev = torch.empty_like(links[0], device='cuda'); dctrl = torch.empty_like(ctrl[0], device='cuda')
for name, fn, nbytes in (('d2h links', lambda: links[0].copy_(dev, non_blocking=True), links[0].numel()*4),
                         ('h2d ctrl', lambda: dctrl.copy_(ctrl[0], non_blocking=True), ctrl[0].numel()*4)):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t)/10
    print(f'{name}: {dt*1e3:.3f} ms, {nbytes/dt/1e9:.1f} GB/s')
def run(tag, up, down, wait_slots=True):
    calls = 0
    pend = [None]*NS
    def one():
        nonlocal calls
        k = calls % NS; calls += 1
        if pend[k] is not None: ph.host_wait_call(pend[k])
        pend[k] = ph.step_host(inner, ctrl=ctrl[k] if up else None, links_row=links[k] if down else None,
                               joints_row=joints[k] if down else None, pipelined=True)
    for _ in range(3): one()
    ph.host_wait(); t = time.perf_counter()
    for _ in range(reps): one()
    ph.host_wait(); dt = (time.perf_counter() - t)/reps
    print(f'{tag}: {dt*1e3:.3f} ms per launch, {n*inner/dt:.4g} env-steps/s')
if not os.environ.get('FARMS_B200_TRACE'):
    run('kernel only (no copies)', False, False)
    run('ctrl up only', True, False)
run('rows down only', False, True)
run('up + down', True, True)
