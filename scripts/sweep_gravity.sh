#!/bin/bash
# Kernel-tuning aid (GPU box): gravity tile-shape / packed-FP32 variants.
python - <<'PY'
import sys; sys.path.insert(0, ".")
from physicl_b200 import _capi
c = _capi.Context(0)
print("FFMA peak %.1f TFLOP/s, FFMA2 (packed) peak %.1f TFLOP/s" % (c.fp32_peak_tflops(), c.fp32x2_peak_tflops()))
PY
for v in ${VARIANTS:-2 10 11 12 13 14 15 0}; do
  PCL_GRAV_VARIANT=$v python bench.py --workload gravity_256k --steps 3 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v: %.2f ms/step %.1f TFLOP/s frac %.3f' % (d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac']))"
done
