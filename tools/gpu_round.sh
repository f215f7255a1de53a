#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, ncu --set full of the three hot encode kernels.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-decode > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lpc4|k_analyze3|k_pack3' --launch-skip 28 -c 4 \
    -o gpurun_out/${tag}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-decode > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
