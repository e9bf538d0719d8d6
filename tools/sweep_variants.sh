#!/bin/bash
# Bench several experiment builds of the library in one GPU session.  Build them first (CPU container):
#   RS_BUILD_TAG=v1 RS_NVCC_EXTRA="-DRS_WSUM_V1" python network-slicing_b200/build.py
# then:  gpurun -- 'bash tools/sweep_variants.sh v1 v2 ...'   ("base" = the default library)
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset RS_B200_LIB; else export RS_B200_LIB=$PWD/network-slicing_b200/libranslice_b200_$tag.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_var_$tag.json 2> gpurun_out/bench_var_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_var_%s.json" % tag))
    print("%-12s %.3fM env-steps/s  %.3f ms/step  kernel %.3f ms" % (tag, d["value"] / 1e6, d["ms_per_step"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print(tag, "FAILED", e)
PY
done
