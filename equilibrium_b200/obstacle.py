"""Rectangle obstacle, mirroring src/simulation/obstacle.rs of the reference."""
from __future__ import annotations

from .configs import SimulationConfigs


class Rectangle:
    """obstacle.rs:29-87.  Axis-parallel rectangle given by two corner points.

    `Rectangle(...)` raises ValueError("Invalid input for Rectangle") where the
    reference panics (obstacle.rs:66-68); the points are public and may be
    mutated afterwards without re-validation, as the reference's GUI does
    (obstacle_widget.rs:176-188).
    """

    def __init__(self, down_left_point, up_right_point, fluid_container_size: int):
        self.down_left_point = (int(down_left_point[0]), int(down_left_point[1]))
        self.up_right_point = (int(up_right_point[0]), int(up_right_point[1]))
        if not self.are_all_points_valid(int(fluid_container_size)):
            raise ValueError("Invalid input for Rectangle")

    @classmethod
    def new(cls, down_left_point, up_right_point, fluid_container_size: int) -> "Rectangle":
        return cls(down_left_point, up_right_point, fluid_container_size)

    @classmethod
    def default(cls) -> "Rectangle":
        """obstacle.rs:47-51: (80,80)-(110,110) on the default 128 grid."""
        return cls((80, 80), (110, 110), SimulationConfigs().size)

    def are_all_points_valid(self, fluid_container_size: int) -> bool:
        """obstacle.rs:74-87"""
        (x0, y0), (x1, y1) = self.down_left_point, self.up_right_point
        return (
            x0 != x1 and y0 != y1 and x0 < x1 and y0 < y1
            and all(e < fluid_container_size for e in (x0, y0, x1, y1))
        )

    def get_approximate_points(self):
        """obstacle.rs:4-7, :83-87: [down_left, up_right]."""
        return [self.down_left_point, self.up_right_point]

    def clone(self) -> "Rectangle":
        r = object.__new__(Rectangle)
        r.down_left_point = self.down_left_point
        r.up_right_point = self.up_right_point
        return r


# obstacle.rs:12-16: `enum ObstaclesType { Rectangle(Rectangle) }` -- in Python
# any object with get_approximate_points() is an obstacle.
ObstaclesType = Rectangle
