#!/bin/bash
# Kernel-tuning aid (GPU box): time the photon kernels with alternative builds of the library.
for lib in "" $EXTRA_LIBS; do
  for tma in 1 0; do
    PCL_PHOTON_TMA=$tma PHYSICL_B200_LIB=$PWD/physicl_b200/libphysicl_b200$lib.so python scripts/tune_photon.py | sed "s/^/tma=$tma /"
  done
done
