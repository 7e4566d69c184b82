set -x
timeout 240 python -m pytest tests/test_trellis_gpu.py -q -x -k "banded" > gpurun_out/r2B_band.log 2>&1; rc=$?; echo "band rc=$rc"; tail -5 gpurun_out/r2B_band.log
if [ $rc -ne 0 ]; then export LNASR_BAND_KERNEL=1; fi
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2B_pytest.log 2>&1; tail -8 gpurun_out/r2B_pytest.log
timeout 120 python tools/probe_umma_rounding.py > gpurun_out/r2B_umma_rounding.txt 2>&1; tail -3 gpurun_out/r2B_umma_rounding.txt
timeout 300 python tools/probe_tc_bias.py > gpurun_out/r2B_tc_bias.txt 2>&1; cat gpurun_out/r2B_tc_bias.txt
timeout 300 python tools/bench_trellis.py > gpurun_out/r2B_trellis.txt 2>&1; cat gpurun_out/r2B_trellis.txt
timeout 600 python tools/e2e_sweep.py c4 > gpurun_out/r2B_e2e_sweep.txt 2>&1; tail -40 gpurun_out/r2B_e2e_sweep.txt
timeout 600 python bench.py --no-cpu > gpurun_out/r2B_bench.log 2> gpurun_out/r2B_bench.err; cut -c1-1500 gpurun_out/r2B_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trellis_band|gmm_tc16_kernel|mfcc_frames" -c 3 -f -o gpurun_out/r2B_c4 python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2B_ncu.log 2>&1; tail -3 gpurun_out/r2B_ncu.log
