import sys, numpy as np
sys.path[:0]=['tests','mamba.jl_b200']
import helpers
from mambacuda.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "pumps_gibbs_amwg"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
tpl, blocks, inits = helpers.scheme(name)
eng = Engine(tpl, C, seed=1)
eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
eng.run(256, burnin=128, thin=10, store=False, out=False)
eng.run(256, burnin=128, thin=10, store=False, out=False)
print(name, C, eng.last_kernel_ms(), "ms per 256 iterations ->", C * 256 / eng.last_kernel_ms() * 1e3, "chain-iterations/s")
