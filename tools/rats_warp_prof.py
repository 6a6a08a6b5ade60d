"""One short launch of the warp-per-chain rats kernel for ncu (65,536 chains x 60 adaptive iterations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests")]
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
eng = Engine(tpl, 65536, seed=123); eng.set_scheme(blocks)
for rep in range(2):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(60, burnin=60, thin=1, store=False, out=False, partial=True)
    print("kernel ms", eng.last_kernel_ms(), "leapfrogs", eng.work_count()[0], flush=True)
