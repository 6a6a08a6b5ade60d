#!/bin/bash
# Times every kernel-variant library under mambacuda/variants/ with a shortened headline bench (experiments only).
cd "$(dirname "$0")/.."
for lib in mamba.jl_b200/mambacuda/variants/lib_*.so; do
  MCU_LIB_PATH=$PWD/$lib timeout 300 python bench.py "$@" --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib'.split('/')[-1], 'kernel_ms', round(r['kernel_ms'],3), 'value', '%.4g'%d['value'], 'frac', round(r['frac'],4), d['clocks']['sm_mhz'], d['config'].get('psrf_max'))" || echo "$lib FAILED"
done
