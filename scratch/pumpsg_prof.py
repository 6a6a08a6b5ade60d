import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mamba.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("pumps_gibbs_amwg")
C = 1000000
eng = Engine(tpl, C, seed=1); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
eng.run(100, burnin=50, thin=10, store=False, out=False)
eng.run(200, burnin=0, thin=10, store=False, out=False)
print("kernel ms", eng.last_kernel_ms(), "chain-it/s", C * 200 / eng.last_kernel_ms() / 1e-3)
