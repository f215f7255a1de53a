#!/bin/bash
# A/B on the GPU box: the quick encode bench line for each variant library (tools/build_variant.sh), then the encode parity
# tests for the variants named after "--".   usage: tools/ab_bench.sh <tag> <variant>... [-- <variant to test>...]
tag=$1; shift
mkdir -p gpurun_out
testing=0
for v in "$@"; do
  if [ "$v" = "--" ]; then testing=1; continue; fi
  lib=flac_codec_b200/libflacb200_$v.so; [ "$v" = "base" ] && lib=flac_codec_b200/libflacb200.so
  if [ $testing = 0 ]; then
    FLACB200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-decode --no-configs --no-files > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
    python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms_per_step", round(d["ms_per_step"], 3), "kernels", {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  else
    FLACB200_LIB=$PWD/$lib python -m pytest tests/test_gpu_encode.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${tag}_test_$v.log 2>&1
    echo "$v tests: $(tail -1 gpurun_out/${tag}_test_$v.log)"
  fi
done
