import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mamba.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, helpers
from mambacuda.engine import Engine
for name in ("dyes_nuts_slice", "dyes_mala_slice", "dyes_hmc_slice"):
    for C, seed in ((16, 123), (16, 7), (1024, 3)):
        tpl, blocks, inits = helpers.scheme(name)
        eng = Engine(tpl, C, seed=seed); eng.set_scheme(blocks); eng.set_inits(inits)
        eng.run(10000, burnin=2500, thin=2, store=False, out=False)
        s = eng.summary_streaming(); nm = eng.names(1)
        print(name, C, seed, {k: (round(s[nm.index(k), 0], 2), round(s[nm.index(k), 3], 2)) for k in ("theta", "mu[5]", "mu[6]", "s2_within")})
