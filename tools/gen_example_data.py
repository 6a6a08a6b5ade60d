"""Prints the data tables of doc/examples/{oxford,epil}.jl as C++ initialiser lists (run in the build container: reads /root/reference).
The tables are pasted into oracle/templates.hpp and mamba.jl_b200/csrc/api.cu (default inputs of the templates), like the other examples' data."""
import re

REF = "/root/reference/doc/examples"


def arr(src, name):
    m = re.search(r":%s =>\s*\[(.*?)\]" % name, src, re.S)
    return [float(v) for v in re.findall(r"-?\d+\.?\d*", m.group(1))]


def fmt(a, per=30):
    s = [("%g" % v) for v in a]
    return ",\n      ".join(", ".join(s[i:i + per]) for i in range(0, len(s), per))


ox = open(f"{REF}/oxford.jl").read()
for k in ("r1", "n1", "r0", "n0", "year"):
    a = arr(ox, k); assert len(a) == 120
    print(f'  in["{k}"] = {{{fmt(a)}}};')
ep = open(f"{REF}/epil.jl").read()
y = arr(ep, "y"); assert len(y) == 236                       # 59 x 4 matrix literal, one patient per row
ycm = [y[i * 4 + j] for j in range(4) for i in range(59)]    # Julia stores it column-major: visit j fastest-varying LAST
print(f'  in["y"] = {{{fmt(ycm)}}};   // 59 x 4, column-major (patient fastest)')
for k in ("Trt", "Base", "Age", "V4"):
    a = arr(ep, k); print(f'  in["{k}"] = {{{fmt(a)}}};')
