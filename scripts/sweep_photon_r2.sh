#!/bin/bash
# photon_sphere_16m / wavelength_64m with alternative builds of the library (PHYSICL_B200_LIB), run under gpurun
for v in "" _minb3 _minb5 _minb6 _blk128 _blk128b; do
  lib=$PWD/physicl_b200/libphysicl_b200$v.so
  [ -f "$lib" ] || continue
  for w in photon_sphere_16m wavelength_64m; do
    r=$(PHYSICL_B200_LIB=$lib python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f G  launches %d  chk %s' % (d['value']/1e9, d['gpu_launches'], d['tally_checksum']))")
    echo "lib${v:-_default} $w : $r"
  done
done
