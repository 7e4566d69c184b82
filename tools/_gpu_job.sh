set -x
timeout 900 python -m pytest tests/test_mfcc_gpu.py tests/test_from_pcm_gpu.py tests/test_guard_bands_gpu.py tests/test_vad_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/job_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/job_pytest.log
