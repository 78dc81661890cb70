"""Frame time of one mode on a bench workload (A/B runs of kernel variants via EQUILIBRIUM_CUDA_LIB).
usage: prof_frame.py [c4|c3|c2] [exact|red_black] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
mode = sys.argv[2] if len(sys.argv) > 2 else "red_black"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
f = bench.build_fluid(wl, mode)
ms = bench.time_device_resident(f, wl["size"], steps, 3, seed=0)
f.profile_reset()
f.profile_enable(True)
f.step_n(steps)
p = f.profile()
print(f"{wl['size']}^2 K={wl['k']} {mode}: {ms / steps:.3f} ms/frame  lin_solve {p['lin_solve_ms'] / steps:.3f} advect {p['advect_ms'] / steps:.3f} "
      f"project {p['project_ms'] / steps:.3f} other {p['other_ms'] / steps:.3f}")
f.close()
