"""Throughput of the example-script schemes that run on the generic one-chain-per-thread kernel (no fused kernel):
65,536 chains x ITERS iterations each, CUDA-event time of the sampler kernel.  python scratch/templates_bench.py > profiles/r1_templates_bench.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mamba.jl_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import helpers
from mambacuda.engine import Engine

C = 65536
rows = []
for name, iters in (("surgical_nuts_slice", 100), ("surgical_amwg", 400), ("dyes_nuts_slice", 100), ("dyes_hmc_slice", 400), ("dyes_mala_slice", 400),
                    ("dyes_rwm_slice", 400), ("salm_slice_amwg", 200), ("equiv_nuts_slice", 100), ("equiv_amwg", 200), ("blocker_amwg_slice", 100),
                    ("stacks_nuts_slice", 100), ("stacks_amwg", 400), ("line_amwg_slice", 1000), ("line_nuts_slice", 200)):
    tpl, blocks, inits = helpers.scheme(name)
    eng = Engine(tpl, C, seed=1)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.02)
    eng.run(iters, burnin=iters // 2, thin=10, store=False, out=False)      # warm-up (adaptation included)
    eng.run(iters, burnin=iters // 2, thin=10, store=False, out=False)
    ms = eng.last_kernel_ms()
    psrf = float(np.nanmax(eng.gelman(0.05, False)[:, 0]))
    rows.append(dict(scheme=name, template=tpl, chains=C, iters=iters, kernel_ms=ms, chain_iters_per_sec=C * iters / (ms * 1e-3)))
    print(name, f"{rows[-1]['chain_iters_per_sec']:.3g}", f"{ms:.1f} ms", file=sys.stderr)
    eng.close()
print(json.dumps(dict(device="1 x B200", kernel="run_generic_kernel<M> (dense instantiation)", rows=rows)))
