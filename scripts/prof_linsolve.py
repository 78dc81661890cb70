"""One exact lin_solve on a config-3-like grid, for ncu / timing experiments.
usage: prof_linsolve.py N K ORIENT [reps] [mode]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from equilibrium_b200 import Fluid, FluidConfigs, SimulationConfigs, Rectangle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity import random_rects

n, k, orient = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mode = sys.argv[5] if len(sys.argv) > 5 else "exact"
f = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), mode=mode)
for r in random_rects(n, 64 if n <= 4096 else 16, n):
    f.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
rng = np.random.default_rng(0)
f.upload("velocities_x", rng.standard_normal((n, n)).astype(np.float32))
f.upload("velocities_x0", rng.standard_normal((n, n)).astype(np.float32))
f.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, 1)   # builds tables, warms up
f.sync()
best = 1e9
for _ in range(reps):
    f.timer_start()
    f.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    best = min(best, f.timer_stop())
f.sync()
cells = (n - 2) ** 2 * k
print(f"N={n} K={k} orient={orient} mode={mode}: {best:.3f} ms  {cells/best/1e6:.1f} Gcell-iter/s  "
      f"{12*cells/best/1e6:.0f} GB/s algorithmic")
