set -x
python bench.py > gpurun_out/r1_v9_bench.json 2> gpurun_out/r1_v9_bench.err
python tools/bench_configs.py > gpurun_out/r1_v9_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r1_v9_configs.json
python tools/md5_bench.py > gpurun_out/r1_v9_md5.json 2>&1
python tools/decode_breakdown.py > gpurun_out/r1_v9_decode_breakdown.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_v9_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_v9_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_lpc3|k_analyze3|k_pack3" -s 6 -c 3 -o gpurun_out/r1_v9_enc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decode > gpurun_out/r1_v9_ncu_enc.log 2>&1
ncu --set full --clock-control none -k regex:"k_parse|k_restore|k_find|k_crc16f|k_chain|k_emit" -c 8 -o gpurun_out/r1_v9_dec -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_v9_ncu_dec.log 2>&1
tail -c 600 gpurun_out/r1_v9_bench.json
# the reports are large: keep the summaries (gpurun_out is capped at 64 MiB)
python tools/ncu_summary.py full gpurun_out/r1_v9_enc.ncu-rep gpurun_out/r1_v9_enc_summary.csv
python tools/ncu_summary.py full gpurun_out/r1_v9_dec.ncu-rep gpurun_out/r1_v9_dec_summary.csv
rm -f gpurun_out/r1_v9_enc.ncu-rep gpurun_out/r1_v9_dec.ncu-rep
