set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r1b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench_c2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_ref.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1b_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_ncu_c2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:mfcc_frames_kernel|mfcc_finalize_kernel|gmm_loglik_exact_kernel|trellis_small_kernel|backtrace_kernel' -s 15 -c 5 -o gpurun_out/r1b_prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_ncufull_c2.log 2>&1
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r1b_bench_c4.log 2>&1
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1b_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_ncu_c4.log 2>&1
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:gmm_tc_kernel|gmm_tc_stage_kernel|trellis_reg_kernel' -s 9 -c 3 -o gpurun_out/r1b_prof_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_ncufull_c4.log 2>&1
nvidia-smi > gpurun_out/r1b_smi.log; lscpu | head -20 >> gpurun_out/r1b_smi.log
