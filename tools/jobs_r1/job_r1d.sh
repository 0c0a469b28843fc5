#!/bin/bash
# Round-1 evidence job (run under gpurun): GPU tests, bench both arms, ncu launch list of the bench
# command, DRAM traffic of the two dominant kernels, --set full captures, phase timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
MAKB200_BQR_NBO=128 timeout 300 python -m pytest tests/test_gpu_qr.py -m gpu -x -q -k "batched" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench6.json 2>> gpurun_out/bench.err
kill $SMI
tail -1 gpurun_out/bench6.json | cut -c1-1500
MAKB200_PROFILE=1 python tools/config_sweep.py C1 C2 > gpurun_out/phases.log 2>&1; grep "makb200 profile" gpurun_out/phases.log | tail -24
# launch list of the bench command (shares, not absolutes)
MAKB200_BENCH_UNDER_NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv \
   --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summ.py gpurun_out/launches2.csv > gpurun_out/launches2_summary.txt; head -30 gpurun_out/launches2_summary.txt
gzip -f gpurun_out/launches2.csv
# DRAM traffic of the two dominant kernels over one whole step
MAKB200_BENCH_UNDER_NCU=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:"gemm_kernel|trd_symv" -c 200000 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
python tools/traffic_summ.py gpurun_out/traffic.csv gpurun_out/traffic.json | tail -30
gzip -f gpurun_out/traffic.csv
# full captures of the dominant kernels
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trd_symv -s 2000 -c 2 -o gpurun_out/r1_symv_full -f python tools/prof_run.py eigh 8192 > /dev/null 2>&1
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 3 -o gpurun_out/r1_gemm_full -f python tools/prof_run.py eigh 8192 > /dev/null 2>&1
# batched QR 257-512 launch breakdown
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mak::" --csv --log-file gpurun_out/bq3_257.csv python tools/batched_prof.py qr 257 512 > gpurun_out/bq3_257.log 2>&1
python tools/ncu_summ.py gpurun_out/bq3_257.csv | head -12; rm -f gpurun_out/bq3_257.csv
ls -la gpurun_out | tail -12
