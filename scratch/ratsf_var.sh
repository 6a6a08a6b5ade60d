for lib in mamba.jl_b200/mambacuda/variants/lib_b*.so; do echo -n "$(basename $lib) "; MCU_LIB_PATH=$PWD/$lib python scratch/generic_prof.py rats_slice_amwg 262144 | tail -1; done
