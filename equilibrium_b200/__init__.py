"""equilibrium_b200: the stable-fluids step of vkabadzhova/equilibrium
(src/simulation/fluid.rs) as hand-written sm_100a CUDA kernels behind the
reference's own `Fluid` API.  Importing this package does not load the native
library; creating a `Fluid` does, and fails loudly if it is missing."""
from .configs import FluidConfigs, SimulationConfigs
from .fluid import ContainerWall, Fluid, connect_distributed, connect_local
from .obstacle import ObstaclesType, Rectangle
from ._lib import EquilibriumError
from .simulation import CurrentSimulation, FluidStep

__all__ = ["Fluid", "FluidConfigs", "SimulationConfigs", "Rectangle", "ObstaclesType",
           "ContainerWall", "EquilibriumError", "CurrentSimulation", "FluidStep", "connect_local", "connect_distributed"]
