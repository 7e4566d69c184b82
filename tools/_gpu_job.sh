set -x
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2D_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2D_pytest.log
timeout 300 python tools/bench_gmm.py > gpurun_out/r2D_gmm.txt 2>&1; cat gpurun_out/r2D_gmm.txt
timeout 600 python bench.py --no-cpu > gpurun_out/r2D_bench.log 2> gpurun_out/r2D_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2D_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['stage_ms'], d['clocks'])
for k,v in d['secondary'].items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['roofline'].get('stage_ms'))
PY
tail -3 gpurun_out/r2D_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mfcc_frames" -c 1 -f -o gpurun_out/r2D_mfcc python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2D_ncu.log 2>&1; tail -2 gpurun_out/r2D_ncu.log
timeout 900 python tools/mfcc_sweep.py > gpurun_out/r2D_mfcc_sweep.md 2>&1; tail -30 gpurun_out/r2D_mfcc_sweep.md
