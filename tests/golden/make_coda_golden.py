"""Writes tests/golden/coda.json: what doc/mcmc/readcoda.jl reads from the OpenBUGS CODA files shipped with the reference
(doc/mcmc/line{1,2}.{out,ind}), parsed here with plain Python (independently of mambacuda.api.readcoda, which the test
checks against this file).  Run in the build container, where /root/reference exists:  python tests/golden/make_coda_golden.py"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
D = "/root/reference/doc/mcmc"


def parse(stem):
    rows = [ln.split() for ln in open(f"{D}/{stem}.out") if ln.strip()]
    it = np.array([int(float(r[0])) for r in rows]); val = np.array([float(r[1]) for r in rows])
    cols, names = [], []
    for ln in open(f"{D}/{stem}.ind"):
        nm, a, b = ln.split()
        names.append(nm); cols.append((it[int(a) - 1:int(b)], val[int(a) - 1:int(b)]))
    assert all((c[0] == cols[0][0]).all() for c in cols)      # every parameter monitored over the same iterations in these files
    return names, cols[0][0], np.stack([c[1] for c in cols], axis=1)


def main():
    n1, it1, v1 = parse("line1"); n2, it2, v2 = parse("line2")
    assert n1 == n2 and (it1 == it2).all()
    v = np.stack([v1, v2], axis=2)
    step = int(it1[1] - it1[0])
    g = {"_about": "doc/mcmc/line{1,2}.{out,ind} parsed as doc/mcmc/readcoda.jl does; see make_coda_golden.py",
         "header": f"Iterations = {it1[0]}:{it1[-1]}\nThinning interval = {step}\nChains = 1,2\nSamples per chain = {len(it1)}\n",
         "names": n1, "column_sums": v.sum(axis=0).tolist(), "head_chain1": v[:3, :, 0].tolist()}
    with open(os.path.join(HERE, "coda.json"), "w") as f:
        json.dump(g, f)
    print("wrote coda.json", v.shape)


if __name__ == "__main__":
    main()
