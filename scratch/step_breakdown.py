import sys, time, numpy as np, torch
sys.path[:0]=['.','tests','mamba.jl_b200']
import bench
from mambacuda import distributed as mdist
from mambacuda.engine import Engine
torch.cuda.set_device(0); device=torch.device("cuda",0)
C=125000
eng = Engine("seeds", C, seed=bench.SEED, device=0); eng.set_scheme(bench.SCHEME)
inits2 = bench.seeds_inits()
l2 = torch.empty(256 << 20, dtype=torch.uint8, device=device)
def t(f):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); return (time.perf_counter()-t0)*1e3, r
for rep in range(4):
    a,_=t(lambda: l2.fill_(1))
    b,_=t(lambda: eng.set_inits(inits2, jitter_sd=0.1))
    c,_=t(lambda: eng.run(2000, burnin=1000, thin=10, store=False, out=False))
    d,_=t(lambda: mdist.global_gelman(eng, 0.05, True, device))
    print(f"flush {a:.2f}  set_inits {b:.2f}  run {c:.2f} (kernel {eng.last_kernel_ms():.2f})  gelman {d:.2f} ms")
import ctypes as Ct
def tt(name, f, n=5):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter()-t0)*1e3/n:.3f} ms"); return r
mm = tt("minmax", lambda: eng.minmax())
codes = tt("link_codes", lambda: eng.link_codes(True, mm))
s0 = tt("moments(None)", lambda: eng.moments(codes, None))
center = np.stack([s0[0][:, 1] / s0[0][:, 0], s0[0][:, 3] / s0[0][:, 0]], axis=1)
s1 = tt("moments(center)", lambda: eng.moments(codes, center))
tt("gelman_from_moments", lambda: eng.gelman_from_moments(s1[1], center, s1[0], 0.05))
tt("summary_streaming", lambda: eng.summary_streaming())
