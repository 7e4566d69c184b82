set -x
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2N_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2N_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2N_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2N_pytest.log
