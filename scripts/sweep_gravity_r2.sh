#!/bin/bash
# gravity_256k: j-split count x kernel variant (run under gpurun)
for v in 0 16 17 18 10; do for ns in 0 2 4 8; do
  r=$(PCL_GRAV_VARIANT=$v PCL_GRAV_NSPLIT=$ns python bench.py --workload gravity_256k --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms  frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))")
  echo "variant $v nsplit $ns : $r"
done; done
