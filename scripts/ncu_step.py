"""One eagerly launched Navier2D.update() inside a cudaProfilerStart/Stop window (for ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rustpde_b200 as R

kind, nx, ny = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
n = R.Navier2D.new_periodic(nx, ny, 1e7, 1.0, 1e-3, 1.0) if kind == "p" else R.Navier2D.new(nx, ny, 1e7, 1.0, 1e-3, 1.0, True)
n.set_graph(False)
n.set_velocity(0.2, 1.0, 1.0); n.set_temperature(0.2, 1.0, 1.0)
n.update(2); n.sync()
torch.cuda.profiler.start()
n.update(steps); n.sync()
torch.cuda.profiler.stop()
print("done", n.div_norm())
