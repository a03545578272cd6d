#!/bin/bash
# Build kernel variants (FMPM_DEFS) next to the product library and time them on the GPU box:
#   CPU container:  bash profiles/ab_variants.sh build "name1:-DG2P_MINB=6 -DG2P_ROUNDS=8" "name2:..."
#   GPU box:        bash profiles/ab_variants.sh run
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
if [ "$1" = build ]; then
  shift
  for spec in "$@"; do
    name="${spec%%:*}"; defs="${spec#*:}"
    FMPM_DEFS="$defs" FMPM_OUT="$PWD/gpurun_variants/$name.so" FMPM_OBJDIR="/tmp/fmpm_obj_$name" python fluidlab_b200/csrc/build.py --force > /dev/null
    echo "built $name ($defs)"
  done
else
  for so in gpurun_variants/*.so; do FMPM_LIB="$PWD/$so" python profiles/phase_times.py 2>&1 | tail -1; done
fi
