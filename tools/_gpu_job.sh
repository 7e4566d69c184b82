set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mfcc_" -c 3 --launch-skip 9 -f -o gpurun_out/r2K_c5_1x1 python tools/mfcc_sweep.py --only 1x1 > gpurun_out/r2K_c5_1x1.log 2>&1; tail -4 gpurun_out/r2K_c5_1x1.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_trellis_gpu.py -q -x -k "banded_decode and 64" > gpurun_out/r2K_memcheck.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/r2K_memcheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trellis_small|gmm_loglik_mma|mfcc_frames" -c 3 -f -o gpurun_out/r2K_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2K_ncu_c2.log 2>&1; tail -2 gpurun_out/r2K_ncu_c2.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2K_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
timeout 600 python tools/e2e_sweep.py c4 > gpurun_out/r2K_e2e_sweep.txt 2>&1; tail -24 gpurun_out/r2K_e2e_sweep.txt
