"""Long runs of the benchmark configurations: diagnostics stay finite and smooth (development aid)."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustpde_b200 as R

def run(periodic, nx, ny, ra, dt, nsteps, every):
    n = R.Navier2D.new_periodic(nx, ny, ra, 1.0, dt, 1.0) if periodic else R.Navier2D.new(nx, ny, ra, 1.0, dt, 1.0, True)
    n.set_velocity(0.2, 1.0, 1.0); n.set_temperature(0.2, 1.0, 1.0)
    t0 = time.time()
    for k in range(0, nsteps, every):
        n.update(every)
        nu, nuvol, re, div, ek = n.eval(True, True, True, True, True)
        print("%s %dx%d step %5d  Nu %.10f  Nuvol %.10f  Re %.8f  |div| %.3e  Ekin %.10e" % ("periodic" if periodic else "confined", nx, ny, k + every, nu, nuvol, re, div, ek), flush=True)
        assert all(math.isfinite(v) for v in (nu, nuvol, re, div, ek))
    print("   %.1f s wall for %d steps" % (time.time() - t0, nsteps), flush=True)

run(False, 2048, 2049, 1e9, 1e-4, 3000, 500)
run(True, 2048, 2049, 1e9, 1e-4, 3000, 500)
run(True, 8192, 8193, 1e10, 2e-5, 300, 100)
