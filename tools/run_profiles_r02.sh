#!/bin/bash
# Round-2 profile pass (run under gpurun): launch list of the bench command, --set full of the encode and decode kernels at the
# bench's launch-group size, the same for the optional fused kernel, executed-opcode histograms (SASS) of the three encode kernels.
tag=${1:-r02_v1}
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-files > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_lpc4|k_analyze3|k_pack3" -s 28 -c 4 -o gpurun_out/${tag}_enc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decode --no-configs > gpurun_out/${tag}_ncu_enc.log 2>&1
ncu --set full --clock-control none -k regex:"k_parse|k_restore|k_find|k_crc16f|k_chain|k_emit" -c 14 -o gpurun_out/${tag}_dec -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-files > gpurun_out/${tag}_ncu_dec.log 2>&1
FLACB200_FUSED=1 ncu --set full --clock-control none -k regex:"k_frame4" -s 2 -c 1 -o gpurun_out/${tag}_fused -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decode --no-configs > gpurun_out/${tag}_ncu_fused.log 2>&1
python tools/ncu_summary.py full gpurun_out/${tag}_enc.ncu-rep gpurun_out/${tag}_enc_summary.csv
python tools/ncu_summary.py full gpurun_out/${tag}_dec.ncu-rep gpurun_out/${tag}_dec_summary.csv
python tools/ncu_summary.py full gpurun_out/${tag}_fused.ncu-rep gpurun_out/${tag}_fused_summary.csv
python tools/ncu_summary.py list gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.md
for k in k_lpc4 k_analyze3 k_pack3; do python tools/sass_hist.py gpurun_out/${tag}_enc.ncu-rep $k 25 > gpurun_out/${tag}_sass_$k.txt 2>&1; done
python tools/sass_hist.py gpurun_out/${tag}_fused.ncu-rep k_frame4 25 > gpurun_out/${tag}_sass_k_frame4.txt 2>&1
rm -f gpurun_out/${tag}_enc.ncu-rep gpurun_out/${tag}_dec.ncu-rep gpurun_out/${tag}_fused.ncu-rep
ls -la gpurun_out/${tag}_*
# single streams (C1 / C3 / C5 shapes): stage clocks and the launch list of the same command
python tools/latency_probe.py > gpurun_out/${tag}_latency.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_latency_launches.csv python tools/latency_probe.py > /dev/null 2>&1
python tools/ncu_summary.py list gpurun_out/${tag}_latency_launches.csv gpurun_out/${tag}_latency_launches.md
