set -x
timeout 900 python -m pytest tests/test_mfcc_gpu.py tests/test_from_pcm_gpu.py tests/test_guard_bands_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/job_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/job_pytest.log
for c in 1x4096 10x64; do timeout 300 python tools/mfcc_sweep.py --only $c 2>&1 | grep "^| [0-9]"; done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mfcc_frames -c 1 python tools/mfcc_sweep.py --only 1x4096 2>&1 | grep -E "gpu__time|inst_exec|issue_active|fp64" | cut -c1-90
