"""One short launch of the fused rats Slice + AMWG kernel for ncu (65,536 chains x 200 iterations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests")]
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
eng = Engine(tpl, 65536, seed=123); eng.set_scheme(blocks)
for rep in range(2):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(200, burnin=100, thin=5, store=False, out=False)
    print("kernel ms", eng.last_kernel_ms(), flush=True)
