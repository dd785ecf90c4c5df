#!/bin/bash
# Kernel-tuning aid (GPU box): time the photon kernels with alternative builds of the library.
for lib in "" _minb3 _minb4 _b128 $EXTRA_LIBS; do
  PHYSICL_B200_LIB=$PWD/physicl_b200/libphysicl_b200$lib.so python scripts/tune_photon.py
done
