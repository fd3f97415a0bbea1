"""farms_mujoco_b200 -- B200-native batched stepping engine for FARMS animats.

Drop-in for the farms_mujoco simulation loop's hot path:
MuJoCo-subset forward dynamics + swimming drag + farms sensor logging for
thousands of independent environments, computed by hand-written sm_100a CUDA
kernels behind a C ABI (include/farms_b200.h).  There is no CPU fallback: the
engine raises if the CUDA library is missing.
"""

__version__ = '0.1.0'
