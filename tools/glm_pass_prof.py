"""A few tensor-core gradient passes of the GLM template at BASELINE.json configs[3] (N = 10^6, d = 100) for ncu: python tools/glm_pass_prof.py [chains]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba.jl_b200")]
import numpy as np
import bench
from mambacuda.engine import Engine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X, y, _ = bench.glm_synthetic(1_000_000, 100, "logit")
eng = Engine("glm", C, seed=1)
eng.set_data("X", X); eng.set_data("y", y); eng.set_data("family", np.array([0.0]))
eng.set_scheme([dict(kind="nuts", nodes=[0])])
beta = 0.1 * np.random.default_rng(5).standard_normal((C, 100))
for r in range(4):
    eng.glm_gradient(beta, impl=1)
    print("pass ms", eng.last_kernel_ms(), flush=True)
