import sys, numpy as np
sys.path[:0]=['oracle','tests','mamba.jl_b200']
import pyoracle, helpers
from mambacuda.engine import Engine
pyoracle.build()
tpl, blocks, inits = helpers.scheme("line_mala")
eng = Engine(tpl, 16, seed=99); eng.set_scheme(blocks); eng.set_inits(inits)
o = pyoracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
og = eng.run(400, burnin=0, thin=1, force_generic=True)
oo, so, to = o.run(16, inits, 400, burnin=0, thin=1, seed=99)
for c in range(16):
    bad = np.where(~np.isclose(og[:, :, c], oo[:, :, c], rtol=1e-8, atol=1e-10).all(axis=1))[0]
    if bad.size:
        i = bad[0]
        print("chain", c, "first diff at", i, "\n  gpu prev", og[i-1,:,c], "\n  gpu", og[i,:,c], "\n  orc", oo[i,:,c])
    else:
        print("chain", c, "identical")
