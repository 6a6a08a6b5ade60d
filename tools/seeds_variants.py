"""Times the fused seeds kernel (resident arm of bench.py: 125,000 chains x 2,000 iterations) for the library given by MCU_LIB_PATH /
MCU_SEEDS_TPC, and checks 64 scattered chains against the oracle.  Usage: python tools/seeds_variants.py [tag]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers, pyoracle
from mambacuda.engine import Engine

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
tpl, blocks, inits = helpers.scheme("seeds_amwg")
C = 125000
eng = Engine(tpl, C, seed=123); eng.set_scheme(blocks)
ms = []
for rep in range(4):
    eng.set_inits(inits, jitter_sd=0.1)
    eng.run(2000, burnin=1000, thin=10, store=False, out=False)
    ms.append(eng.last_kernel_ms())
best = min(ms[1:])
# parity spot check: 64 scattered chains, 400 iterations
eng.set_inits(inits, jitter_sd=0.1)
out = eng.run(400, burnin=200, thin=10)
st, tune, _ = eng.get_state()
ids = np.unique(np.concatenate([[0, 1, 95, 96, 97, C - 1], np.random.default_rng(0).integers(0, C, 58)]))
orc = pyoracle.Oracle(tpl); orc.set_scheme([helpers.oracle_block(b) for b in blocks])
oo, so, to, marg = orc.run(0, inits, 400, burnin=200, thin=10, seed=123, jitter_sd=0.1, chain_ids=ids, nthreads=os.cpu_count(), margins=True)
kept = [i for i in range(1, 401) if i > 200 and (i - 200) % 10 == 0]
try:
    n_same, ties = helpers.audit_divergence((out[:, :, ids], st[ids], tune[ids]), (oo, so, to), marg, kept, 0)
    par = f"parity {n_same}/{len(ids)} ties {len(ties)}"
except AssertionError as e:
    par = "PARITY FAIL " + str(e)[:200]
print(f"{tag}: kernel {best:.2f} ms  -> {C * 2000 / best / 1e6:.1f}e9 chain-it/s  frac {6026 * C * 2000 / (best * 1e-3) / 34.2e12:.3f}  all {['%.1f' % m for m in ms]}  {par}", flush=True)
