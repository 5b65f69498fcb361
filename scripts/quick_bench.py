"""Ad-hoc timing of Navier2D.update() (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rustpde_b200 as R

def run(periodic, nx, ny, steps=20, warm=5):
    t0 = time.time()
    n = (R.Navier2D.new_periodic(nx, ny, 1e7, 1.0, 1e-3, 1.0) if periodic else R.Navier2D.new(nx, ny, 1e7, 1.0, 1e-3, 1.0, True))
    n.set_velocity(0.2, 1.0, 1.0); n.set_temperature(0.2, 1.0, 1.0)
    t1 = time.time()
    n.update(warm); n.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); n.update(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%s %dx%d: setup %.1fs, %.3f ms/step, %.1f steps/s, launches/step %d, div %.3e" % ("periodic" if periodic else "confined", nx, ny, t1 - t0, ms, 1e3 / ms, n.launches_per_step(), n.div_norm()), flush=True)

def prof(periodic, nx, ny, reps=5):
    n = (R.Navier2D.new_periodic(nx, ny, 1e7, 1.0, 1e-3, 1.0) if periodic else R.Navier2D.new(nx, ny, 1e7, 1.0, 1e-3, 1.0, True))
    n.set_velocity(0.2, 1.0, 1.0); n.set_temperature(0.2, 1.0, 1.0)
    n.update(3); n.sync()
    p = n.profile(reps)
    tot = sum(o["ms"] for o in p)
    print("profile %s %dx%d total %.3f ms" % ("periodic" if periodic else "confined", nx, ny, tot))
    for o in p:
        gbs = o["bytes"] / (o["ms"] * 1e-3) / 1e9 if o["ms"] > 0 else 0
        tf = o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] > 0 else 0
        print("  %-28s %8.3f ms  %5.1f%%  %8.1f GB/s  %6.2f TF/s" % (o["name"], o["ms"], 100 * o["ms"] / tot, gbs, tf))
    sys.stdout.flush()


if __name__ == "__main__":
    torch.cuda.init()
    for a in sys.argv[1:]:
        p, nx, ny = a.split(",")
        if p in ("P", "C"):
            prof(p == "P", int(nx), int(ny))
        else:
            run(p == "p", int(nx), int(ny))

