"""Where does the streaming e2e go?  Times, per step at 2048x2049 confined: H2D only, D2H only, both, both + update."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rustpde_b200 as R
from rustpde_b200 import _ffi
lib = _ffi.product_lib(0)
nav = R.Navier2D.new(2048, 2049, 1e9, 1.0, 1e-4, 1.0, True, lib=lib)
nav.set_velocity(0.2, 1, 1); nav.set_temperature(0.2, 1, 1)
nav.update(3); nav.sync()
fields = [nav.temp, nav.ux, nav.uy, nav.pres[0]]
pin = []
for f in fields:
    a = f.vhat
    t = torch.empty(a.size, dtype=torch.float64).pin_memory(); t.numpy()[:] = a.reshape(-1); pin.append(t)
pout = [torch.empty_like(p).pin_memory() for p in pin]
ptrs = lambda ts: [(t.data_ptr(), t.numel()) for t in ts]
N = 20
def run(name, fn, fin=None):
    fn(); 
    if fin: fin()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): fn()
    if fin: fin()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
    print("%-40s %.3f ms/step" % (name, dt * 1e3), flush=True)
def h2d():
    nav.stage_state(*ptrs(pin)); nav.commit_staged(); nav.sync()
def d2h():
    nav.fetch_state(*ptrs(pout)); nav.fetch_wait()
def both():
    nav.commit_staged(); nav.fetch_state(*ptrs(pout)); nav.stage_state(*ptrs(pin)); nav.sync()
def both_upd():
    nav.commit_staged(); nav.update(1); nav.fetch_state(*ptrs(pout)); nav.stage_state(*ptrs(pin)); nav.sync()
def both_upd_div():
    nav.commit_staged(); nav.update(1); nav.fetch_state(*ptrs(pout)); nav.stage_state(*ptrs(pin)); nav.div_norm()
run("H2D 134 MB (stage+commit)", h2d)
run("D2H 134 MB (fetch+wait)", d2h)
nav.stage_state(*ptrs(pin))
run("H2D + D2H concurrently", both, nav.fetch_wait)
run("H2D + D2H + update", both_upd, nav.fetch_wait)
run("H2D + D2H + update + div_norm", both_upd_div, nav.fetch_wait)
nav.commit_staged()
run("update only", lambda: (nav.update(1), nav.sync()))
# plain torch copies of the same size for comparison (linear, not pitched)
big = torch.empty(134_000_000 // 8, dtype=torch.float64).pin_memory()
dev = torch.empty_like(big, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t_h2d():
    dev.copy_(big, non_blocking=True); torch.cuda.synchronize()
def t_both():
    with torch.cuda.stream(s1): dev.copy_(big, non_blocking=True)
    with torch.cuda.stream(s2): big2.copy_(dev2, non_blocking=True)
    torch.cuda.synchronize()
big2 = torch.empty_like(big).pin_memory(); dev2 = torch.empty_like(dev)
run("torch linear H2D 134 MB", t_h2d)
run("torch linear H2D + D2H concurrently", t_both)
