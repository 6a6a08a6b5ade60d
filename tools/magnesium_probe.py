"""ms per iteration of the generic kernel on the magnesium template, block by block (why is the full scheme slow?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers
from mambacuda.engine import Engine
tpl, blocks, inits = helpers.scheme("magnesium")
for name, bl in [("all", blocks)] + [(f"block{i}:{b['kind']}{b['nodes']}", [b]) for i, b in enumerate(blocks)]:
    for C in (64, 16384):
        eng = Engine(tpl, C, seed=1); eng.set_scheme(bl); eng.set_inits(inits, jitter_sd=0.02)
        eng.run(20, burnin=10, thin=1, store=False, out=False)
        eng.run(200, burnin=10, thin=1, store=False, out=False)
        print(f"{name} C={C}: {eng.last_kernel_ms() / 200:.3f} ms/iteration", flush=True)
        eng.close()
