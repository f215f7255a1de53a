set -x
python bench.py > gpurun_out/r1_v6_bench.json 2> gpurun_out/r1_v6_bench.err
python tools/bench_configs.py > gpurun_out/r1_v6_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r1_v6_configs.json
python tools/md5_bench.py > gpurun_out/r1_v6_md5.json 2>&1
python tools/decode_breakdown.py > gpurun_out/r1_v6_decode_breakdown.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_v6_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_v6_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_lpc3|k_analyze3|k_pack3" -s 6 -c 3 -o gpurun_out/r1_v6_enc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decode > gpurun_out/r1_v6_ncu_enc.log 2>&1
ncu --set full --clock-control none -k regex:"k_parse|k_restore|k_find|k_crc16f|k_chain|k_emit" -c 7 -o gpurun_out/r1_v6_dec -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_v6_ncu_dec.log 2>&1
tail -c 600 gpurun_out/r1_v6_bench.json
