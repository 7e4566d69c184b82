set -x
timeout 900 python -m pytest tests/test_mfcc_gpu.py tests/test_pipeline_gpu.py tests/test_from_pcm_gpu.py tests/test_full_size_gpu.py tests/test_guard_bands_gpu.py -q > gpurun_out/r2M_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2M_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2M_bench.log 2> gpurun_out/r2M_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2M_bench.log') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['roofline']['stage_ms'], d['clocks']['sm_mhz'])
for k,v in d['secondary'].items(): print(k, v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['roofline'].get('stage_ms'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:mfcc_frames -c 2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-secondary 2>&1 | grep -E "gpu__time|inst_executed" | head -4
