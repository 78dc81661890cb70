"""Plain-data configs, mirroring src/simulation/configs.rs of the reference."""
from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass
class SimulationConfigs:
    """configs.rs:5-32 (`SimulationConfigs`, defaults :14-22)."""

    delta_t: float = 0.02
    frames: int = 16
    size: int = 128

    @classmethod
    def new(cls, delta_t: float, frames: int, fluid_container_size: int) -> "SimulationConfigs":
        return cls(delta_t, frames, fluid_container_size)

    def copy(self) -> "SimulationConfigs":
        return replace(self)


@dataclass
class FluidConfigs:
    """configs.rs:37-60 (`FluidConfigs`).  The reference spells it `viscousity`.

    Colours are carried for API compatibility only; they are used by the
    renderer (renderer_helpers.rs:122-167), never by the solver.
    """

    diffusion: float = 0.0
    viscousity: float = 0.001
    has_perlin_noise: bool = True
    fluid_color: tuple = (208, 88, 157, 220)
    world_color: tuple = (94, 146, 162, 128)

    def copy(self) -> "FluidConfigs":
        return replace(self)
