# End-of-round evidence (one GPU).  Every ncu run follows a plain run of the same command that exited 0; numbers printed under
# ncu are never bench values.  Outputs land in gpurun_out/fin_*; the large bench captures are condensed ON THE BOX
# (tools/ncu_summary.py, tools/ncu_traffic.py) because gpurun brings back at most 64 MiB; the per-kernel captures (with source
# pages) come back whole and are condensed on the authoring side (tools/ncu_hotspots.py).
#   bash tools/gpu_round2_final.sh [tests] [bench] [captures] [kernels]
set -x
K='regex:mfcc_frames_kernel|gmm_tc16_kernel|trellis_band2_kernel|gmm_loglik_mma|trellis_small_kernel'
want() { case " $ARGS " in *" $1 "*) return 0;; esac; return 1; }
ARGS="${*:-tests bench captures kernels}"
if want tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/fin_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fin_smoke.log 2>&1; tail -1 gpurun_out/fin_smoke.log
fi
if want bench; then
  timeout 900 python bench.py > gpurun_out/fin_bench_c4.json 2> gpurun_out/fin_bench_c4.err; echo "bench rc=$?"
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.json 2>&1; echo "ref rc=$?"
  timeout 900 python tools/mfcc_sweep.py > gpurun_out/fin_mfcc_sweep.md 2> gpurun_out/fin_mfcc_sweep.err
  timeout 300 python tools/bench_em_large.py 200 4 > gpurun_out/fin_em_large.log 2>&1; timeout 300 python tools/bench_em_large.py 2000 2 >> gpurun_out/fin_em_large.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/fin_launches_em.csv python tools/bench_em_large.py 200 4 > /dev/null 2>&1
fi
if want captures; then
  for w in c2 c4; do
    timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/fin_plain_$w.log 2>&1 &&
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/fin_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/fin_ncu_$w.log 2>&1
    timeout 900 ncu --set full --clock-control none -k "$K" -s 9 -c 3 -f -o gpurun_out/fin_$w python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/fin_ncufull_$w.log 2>&1
    frames=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/fin_plain_$w.log') if l.startswith('{')][-1])['config']['frames_per_gpu_step'])")
    python tools/ncu_summary.py gpurun_out/fin_$w.ncu-rep > gpurun_out/fin_${w}_ncu_full.txt
    python tools/ncu_traffic.py $w $frames gpurun_out/fin_$w.ncu-rep && cp profiles/ncu_traffic.json gpurun_out/fin_ncu_traffic.json
    rm -f gpurun_out/fin_$w.ncu-rep
  done
fi
if want kernels; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_frames -c 1 -f -o gpurun_out/fin_mfcc python tools/mfcc_sweep.py --only 1x4096 > gpurun_out/fin_ncufull_mfcc.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:mfcc -s 3 -c 3 -f -o gpurun_out/fin_mfcc_b1 python tools/mfcc_sweep.py --only 1x1 > gpurun_out/fin_ncufull_mfcc_b1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_moments_mma -c 1 -f -o gpurun_out/fin_emm python tools/bench_em_large.py 200 4 > gpurun_out/fin_ncufull_emm.log 2>&1
fi
du -sh gpurun_out; ls -la gpurun_out/fin_*
