#!/bin/bash
# build_variant.sh NAME "-DFLAGS for hw_ma_tpe.cu" ["-DFLAGS for hw_sa.cu"]: an experiment library in build/variants/NAME.so (other objects reused)
mkdir -p build/variants
HW_BUILD_REUSE_OBJECTS=1 HW_NVCC_EXTRA_HW_MA_TPE="$2" HW_NVCC_EXTRA_HW_SA="$3" python -c "
from gym_highway_b200 import build
print(build.build(force=True, out='/root/repo/build/variants/$1.so'))"
cuobjdump -res-usage build/variants/$1.so 2>/dev/null | grep -A1 "ma_tpe_step_kernelILi5ELb0ELb0\|sa_step_kernelILb0ELb0ELb0" | grep -o "REG:[0-9]* STACK:[0-9]*"
