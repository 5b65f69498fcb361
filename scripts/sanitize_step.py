"""A few steps of every schedule variant on small grids, for `compute-sanitizer --tool memcheck python scripts/sanitize_step.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustpde_b200 as R

def run(periodic, nx, ny, env):
    for k in ("RUSTPDE_B200_XW", "RUSTPDE_B200_XS", "RUSTPDE_B200_PW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    n = R.Navier2D.new_periodic(nx, ny, 1e5, 1.0, 0.01, 1.0) if periodic else R.Navier2D.new(nx, ny, 1e5, 1.0, 0.01, 1.0, True)
    n.set_velocity(0.2, 1.0, 1.0); n.set_temperature(0.2, 1.0, 1.0)
    n.update(3); n.sync()
    print(periodic, nx, ny, env, n.kernel_path(), n.launches_per_step(), "div %.3e" % n.div_norm(), flush=True)

run(False, 72, 65, {})                          # Bluestein x, warp-serial ADI-x, ragged strips (63 columns)
run(False, 72, 65, {"RUSTPDE_B200_XW": "2"})    # + xw_div, xw_project
run(False, 72, 65, {"RUSTPDE_B200_XW": "0"})    # tile kernels only
run(False, 65, 65, {})                          # power-of-two period along x
run(False, 129, 33, {"RUSTPDE_B200_XW": "2"})
run(True, 64, 65, {"RUSTPDE_B200_PW": "1"})     # periodic, row sweeps
run(True, 72 // 8 * 8, 129, {"RUSTPDE_B200_PW": "0"})
