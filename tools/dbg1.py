import numpy as np, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import make_case
from farms_mujoco_b200.engine import BatchedPhysics
spec, model, qpos0, qvel0, ctrl = make_case('salamander_swim', 16)
ref=None
for team in (32,16,8):
  for chunks in ([20],[20],[5,5,5,5],[1]*20,[10,10]):
    ph = BatchedPhysics.from_spec(spec, 16, buffer_size=21, team_lanes=team)
    ph.reset(qpos0, qvel0); ph.set_ctrl(ctrl)
    for n in chunks:
        ph.step(n)
    q=ph.qpos
    if ref is None: ref=q
    print(team, chunks[:3], 'flags', ph.flags.tolist(), 'nan envs', np.where(np.isnan(q).any(axis=1))[0].tolist(), 'maxdiff', np.nanmax(np.abs(q-ref)))
