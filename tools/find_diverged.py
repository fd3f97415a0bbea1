"""Which environments of a bench-like run go non-finite, and when (GPU box)."""
import sys, json
import numpy as np
sys.path.insert(0, '.')
from farms_mujoco_b200 import models, mjcf_subset
from farms_mujoco_b200.engine import BatchedPhysics
from farms_mujoco_b200.sharding import synthetic_inputs
from farms_mujoco_b200.models import travelling_wave_parameters
name, n = sys.argv[1], int(sys.argv[2])
spec = models.MODELS[name](); model = mjcf_subset.parse_mjcf(spec.mjcf)
qpos0, qvel0, phase = synthetic_inputs(model, np.arange(n))
ph = BatchedPhysics.from_spec(spec, n, buffer_size=32)
joints, amp, freq, lag = travelling_wave_parameters(spec)
ph.set_env_phase(phase)
ph.set_wave_controller([model.actuator_id(f'actuator_position_{j}') for j in joints], amp, freq, lag)
ph.reset(qpos0, qvel0)
first = {}
for launch in range(12):
    ph.step(16)
    bad = np.flatnonzero(ph.flags & 1)
    for e in bad:
        first.setdefault(int(e), launch)
print(json.dumps({'n_bad': len(first), 'first': dict(list(first.items())[:12])}))
