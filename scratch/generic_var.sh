run() { echo -n "$1 "; MCU_LIB_PATH=$2 python scratch/generic_prof.py $3 $4 | tail -1; }
for t in line seeds pumps; do
  case $t in line) s=line_amwg_slice;; seeds) s=seeds_amm;; pumps) s=pumps_gibbs_amwg;; esac
  for lib in mamba.jl_b200/mambacuda/variants/lib_${t}_m*.so; do run $(basename $lib) $PWD/$lib $s 262144; done
done
