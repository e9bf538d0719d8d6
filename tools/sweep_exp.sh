#!/bin/bash
# On the GPU box: timing experiments with parts of the eMBB kernel disabled (-DRS_EXP=mask).
for e in "$@"; do
  RS_NVCC_EXTRA="-DRS_EXP=$e" python network-slicing_b200/build.py --force >/dev/null 2>&1
  echo -n "RS_EXP=$e: "
  python bench.py --steps 6 --warmup 3 --burn-in 200 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
