"""One launch of the fused seeds kernel for ncu: 125,000 chains x 300 iterations (or argv[1]) (MCU_SEEDS_TPC / MCU_LIB_PATH select the kernel)."""
import os, sys
ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests")]
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("seeds_amwg")
eng = Engine(tpl, 125000, seed=123); eng.set_scheme(blocks)
for rep in range(2):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(ITERS, burnin=ITERS // 2, thin=10, store=False, out=False)
print("kernel ms", eng.last_kernel_ms())
