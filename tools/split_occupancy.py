from farms_mujoco_b200 import models
from farms_mujoco_b200.engine import BatchedPhysics
spec=models.MODELS["salamander_swim"]()
ph=BatchedPhysics.from_spec(spec, 4096, buffer_size=4)
ph.step(2); ph.step(2)
print("split warps", ph.fast_split, "blocks/SM", ph.lib.fb_fast_split_blocks_per_sm(ph._handle))
