import sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "mamba.jl_b200")
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = Engine(tpl, C, seed=1)
eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
tot = 0
for (n, b) in [(10, 5), (90, 200), (100, 200)]:
    eng.run(n, burnin=b, thin=1, store=False, out=False)
    ms = eng.last_kernel_ms()
    st, tune, it = eng.get_state()
    tot += n
    print(f"iters {tot:4d}: kernel {ms:8.1f} ms  ({ms/n:7.2f} ms/iter)  eps median {np.median(tune[:,2]):.4g}  nalpha(last doubling) mean {tune[:,7].mean():.1f} max {tune[:,7].max():.0f}  alpha/nalpha {np.mean(tune[:,1]/tune[:,7]):.3f}")
