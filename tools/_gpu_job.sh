set -x
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_full_size_gpu.py tests/test_tc_gpu.py tests/test_recognizer_surface.py -q > gpurun_out/r2I_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2I_pytest.log
for i in 1 2 3; do
timeout 600 python bench.py --no-cpu > gpurun_out/r2I_bench$i.log 2> gpurun_out/r2I_bench.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2I_bench$i.log') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'sync', d['e2e']['one_batch_at_a_time']['ms_per_step'], d['roofline']['stage_ms'], d['clocks']['sm_mhz'])
for k,v in d['secondary'].items(): print(k, v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'])
PY
done
tail -3 gpurun_out/r2I_bench.err
