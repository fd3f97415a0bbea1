"""Option objects (farms_core ``AnimatOptions`` / ``ArenaOptions`` /
``SimulationOptions`` stand-ins).

Only the fields the hot path reads are provided, under the reference's field
names (SURVEY.md section 5 "Config / flags"; farms_mujoco/swimming/drag.pyx:
338-385, farms_mujoco/simulation/task.py:274-286, simulation.py:52-79).
"""

from dataclasses import dataclass, field
from typing import List, Optional

from .units import SimulationUnitScaling


@dataclass
class LinkOptions:
    """``animat_options.morphology.links[i]`` (drag.pyx:353-385, mjcf.py:1409-1424)."""
    name: str
    swimming: bool = False
    density: float = 1000.0
    drag_coefficients: List[List[float]] = field(
        default_factory=lambda: [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    friction: List[float] = field(default_factory=lambda: [1.0, 0.0, 0.0])
    mass_multiplier: float = 1.0


@dataclass
class JointOptions:
    """``animat_options.morphology.joints[i]`` (mjcf.py:745-780, 1427-1444)."""
    name: str
    initial: List[float] = field(default_factory=lambda: [0.0, 0.0])
    stiffness: float = 0.0
    damping: float = 0.0
    limits: Optional[List[float]] = None
    extras: dict = field(default_factory=dict)


@dataclass
class MotorOptions:
    """``animat_options.control.motors[i]`` (mjcf.py:798-866, task.py:274-286)."""
    joint_name: str
    control_types: List[str] = field(default_factory=lambda: ['position'])
    gains: List[float] = field(default_factory=lambda: [1.0, 0.0])
    limits_torque: Optional[List[float]] = None

    def __getitem__(self, key):  # the reference indexes motors like dicts (task.py:276)
        return getattr(self, key)


@dataclass
class MorphologyOptions:
    links: List[LinkOptions] = field(default_factory=list)
    joints: List[JointOptions] = field(default_factory=list)
    self_collisions: List[List[str]] = field(default_factory=list)

    def links_names(self):
        return [link.name for link in self.links]

    def joints_names(self):
        return [joint.name for joint in self.joints]


@dataclass
class ControlOptions:
    motors: List[MotorOptions] = field(default_factory=list)
    hill_muscles: list = field(default_factory=list)
    muscles: Optional[list] = None

    def joints_names(self):
        return [motor.joint_name for motor in self.motors]


@dataclass
class SpawnOptions:
    pose: List[float] = field(default_factory=lambda: [0.0]*6)
    velocity: List[float] = field(default_factory=lambda: [0.0]*6)


@dataclass
class AnimatOptions:
    name: str = 'animat'
    sdf: str = ''
    spawn: SpawnOptions = field(default_factory=SpawnOptions)
    morphology: MorphologyOptions = field(default_factory=MorphologyOptions)
    control: ControlOptions = field(default_factory=ControlOptions)
    mujoco: dict = field(default_factory=dict)


@dataclass
class WaterOptions:
    """``arena_options.water`` (drag.pyx:338-350, mjcf.py:1213-1225)."""
    height: Optional[float] = None
    sdf: str = ''
    drag: bool = False
    sph: bool = False
    buoyancy: bool = False
    density: float = 1000.0
    velocity: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    viscosity: float = 1.0


@dataclass
class ArenaOptions:
    sdf: str = ''
    spawn: SpawnOptions = field(default_factory=SpawnOptions)
    ground_height: Optional[float] = None
    water: WaterOptions = field(default_factory=WaterOptions)


@dataclass
class SimulationOptions:
    """Fields read at simulation.py:52,62,76-79,136,141,151 and mjcf.py:1326-1403."""
    timestep: float = 1e-3
    n_iterations: int = 1000
    num_sub_steps: int = 1
    buffer_size: int = 0          # 0 -> n_iterations (reference default use)
    play: bool = True
    headless: bool = True
    fast: bool = True
    show_progress: bool = False
    units: SimulationUnitScaling = field(default_factory=SimulationUnitScaling)
    gravity: List[float] = field(default_factory=lambda: [0.0, 0.0, -9.81])
    impratio: float = 1.0
    cone: str = 'pyramidal'
    solver: str = 'Newton'
    n_solver_iters: int = 100
    integrator: str = 'Euler'
    video: bool = False
