#!/bin/bash
# A/B: current build vs tmp_ab/libnbody_old.so on the same GPU (force and force+potential, plummer and spline)
for lib in NEW OLD; do
  echo "== $lib"
  if [ $lib = OLD ]; then export NBODY_SM100_LIB=tmp_ab/libnbody_old.so; else unset NBODY_SM100_LIB; fi
  python tools/gpu_probe.py --n 262144 --variants 0 2>/dev/null | grep -E "^v0 (plummer|spline) +float32_kahan +(force|both)"
done
