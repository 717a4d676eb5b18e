#!/bin/bash
# One gpurun call of session 4: new strided tests + the parity suites, the bench line, the ncu launch list of the
# bench command and a --set full capture of the three dominant kernels.  Outputs under gpurun_out/s4_*.
mkdir -p gpurun_out
date +%s > gpurun_out/s4_t0
timeout 300 python -m pytest tests/test_gpu_strided.py tests/test_gpu_parity.py tests/test_gpu_extensions.py -q -m gpu > gpurun_out/s4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
timeout 420 python bench.py > gpurun_out/s4_bench_f64.json 2> gpurun_out/s4_bench_f64.err
echo "bench rc=$?" >> gpurun_out/s4_bench_f64.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-records > gpurun_out/s4_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_interpolate_pair2_f64|k_deposit_tiles|k_z_eo' \
    --launch-skip 3 -c 3 -o gpurun_out/s4_top3 python tools/profile_step.py large f64 2 > gpurun_out/s4_ncu_full.log 2>&1
timeout 150 python bench.py --dtype f32 --no-cpu-baseline --no-gpu-baseline > gpurun_out/s4_bench_f32.json 2> gpurun_out/s4_bench_f32.err
timeout 100 python tools/benchmark_sweep.py --no-cpu --deposit > gpurun_out/s4_sweep.log 2>&1
date +%s > gpurun_out/s4_t1
tail -3 gpurun_out/s4_pytest.log; cut -c1-600 gpurun_out/s4_bench_f64.json
