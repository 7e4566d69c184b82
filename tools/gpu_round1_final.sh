# End-of-round evidence, part A: launch lists (ncu --metrics gpu__time_duration.sum) of each workload; every ncu run follows
# a plain run of the same command that exited 0.  Part B (full captures) is run per workload: gpurun brings back <= 64 MiB.
set -x
for w in c2 c3 c4; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > gpurun_out/fin_plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/fin_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > gpurun_out/fin_ncu_$w.log 2>&1
done
