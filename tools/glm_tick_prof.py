"""A short GLM / NUTS run through the tick engine at BASELINE.json configs[3] (N = 10^6, d = 100) for an ncu launch list:
python tools/glm_tick_prof.py [chains] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba.jl_b200")]
import numpy as np
import bench
from mambacuda.engine import Engine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
X, y, _ = bench.glm_synthetic(1_000_000, 100, "logit")
eng = Engine("glm", C, seed=1)
eng.set_data("X", X); eng.set_data("y", y); eng.set_data("family", np.array([0.0]))
eng.set_scheme([dict(kind="nuts", nodes=[0])])
eng.set_inits(np.zeros((1, 100)), jitter_sd=0.1)
eng.run(4, burnin=1, thin=1, store=False, out=False)      # one-time setup (buffers, X pre-pack, X'y on the host) outside the timed call
w0 = eng.work_count(); s0 = eng.glm_pass_slots
eng.set_inits(np.zeros((1, 100)), jitter_sd=0.1)
t0 = time.perf_counter()
eng.run(iters, burnin=iters - 2, thin=1, store=False, out=False)
dt = time.perf_counter() - t0
w = eng.work_count()
ticks = w[1] - w0[1]
print(f"{C} chains x {iters} iterations: {dt * 1e3:.1f} ms, ticks {ticks}, {dt * 1e3 / max(ticks, 1):.3f} ms per tick, pass slots {eng.glm_pass_slots - s0}", flush=True)
