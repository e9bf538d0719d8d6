#!/bin/bash
# On the GPU box: rebuild the library with different occupancy targets and bench each (scratch experiment).
for mb in "$@"; do
  RS_NVCC_EXTRA="-DRS_FAST_MIN_BLOCKS=$mb" python network-slicing_b200/build.py --force >/dev/null 2>&1
  echo -n "min_blocks=$mb: "
  python bench.py --steps 10 --warmup 3 --burn-in 300 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
done
