#!/bin/bash
# sweep kernel variants (different .so files) with the stall probe
for lib in "" scripts/variants/lib_split5.so scripts/variants/lib_split7.so; do
  echo "=== lib: ${lib:-default(split6)}"
  COLPALI_B200_LIB=$lib timeout 200 python scripts/dev_probe2_maxsim.py 2>&1 | grep -E "C=2 (normal|noEpi |noTMA\+noEpi)"
done
