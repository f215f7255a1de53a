#!/bin/bash
# Round-end GPU visit: the whole -m gpu suite, the bench line, then the profile pass (tools/run_profiles_r02.sh).  usage: tools/gpu_final.sh <tag>
tag=${1:-r02_v4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench.json
bash tools/run_profiles_r02.sh $tag > gpurun_out/${tag}_profiles.log 2>&1
cat gpurun_out/${tag}_enc_summary.csv gpurun_out/${tag}_dec_summary.csv | awk 'NR<=2 || !/^Kernel Name|^,,,/' > gpurun_out/${tag}_ncu_full_summary.csv
ls gpurun_out/${tag}_*
