"""Times the warp-per-chain rats kernel (NUTS + Slice; bench.py's rats_slice_amwg leg: 65,536 chains x 2,000 iterations).  Usage: python tools/rats_warp_time.py [tag]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests")]
import helpers
from mambacuda.engine import Engine
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
eng = Engine(tpl, 65536, seed=123); eng.set_scheme(blocks)
ms = []
for rep in range(4):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(2000, burnin=1000, thin=5, store=False, out=False)
    ms.append(eng.last_kernel_ms())
w = eng.work_count()
print(f"{tag}: kernel {min(ms):.0f} ms  all {['%.0f' % m for m in ms]}  leapfrogs {w[0]:.3e}", flush=True)
