set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trellis_band|gmm_tc16_kernel|mfcc_frames" -c 3 -f -o gpurun_out/r2J_c4 python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2J_ncu.log 2>&1; tail -2 gpurun_out/r2J_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2J_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2J_ncu2.log 2>&1; tail -2 gpurun_out/r2J_ncu2.log
