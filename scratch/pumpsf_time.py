import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mamba.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("pumps_slice")
for C in (1000000,):
    eng = Engine(tpl, C, seed=1); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(20, burnin=10, thin=1, store=False, out=False)
    eng.set_inits(inits, jitter_sd=0.05)
    eng.run(500, burnin=250, thin=10, store=False, out=False)
    ms = eng.last_kernel_ms(); print(os.environ.get("MCU_LIB_PATH", "default").split("/")[-1], C, f"{C * 500 / ms / 1e3:.4g}")
    eng.close()
