"""One launch of the fused pumps Gibbs + AMWG kernel for ncu / timing: python tools/pumps_gibbs_prof.py [chains] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests")]
import helpers
from mambacuda.engine import Engine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
tpl, blocks, inits = helpers.scheme("pumps_gibbs_amwg")
eng = Engine(tpl, C, seed=123); eng.set_scheme(blocks)
for rep in range(3):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(iters, burnin=iters // 2, thin=10, store=False, out=False)
    print("kernel ms", eng.last_kernel_ms(), "-> %.3e chain-iterations/s" % (C * iters / eng.last_kernel_ms() * 1e3), flush=True)
