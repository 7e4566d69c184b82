set -x
timeout 900 python -m pytest tests/test_em_gpu.py tests/test_seg_gpu.py -x -q > gpurun_out/job_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/job_pytest.log
timeout 300 python tools/bench_em_large.py 200 4 | tail -1
timeout 300 python tools/bench_em_large.py 2000 2 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:em_moments_mma -c 1 python tools/bench_em_large.py 200 4 2>&1 | grep -E "gpu__time|inst_exec|wavefronts|tensor"
