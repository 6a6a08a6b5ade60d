import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "mamba.jl_b200")
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
eng = Engine(tpl, 18944, seed=1)
eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
eng.run(60, burnin=30, thin=1, store=False, out=False)
eng.run(10, burnin=30, thin=1, store=False, out=False)
print(eng.last_kernel_ms())
