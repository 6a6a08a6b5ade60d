for lib in mamba.jl_b200/mambacuda/variants/lib_*.so; do echo $lib; MCU_LIB_PATH=$PWD/$lib python scratch/rats_probe.py 65536 | tail -1; done
