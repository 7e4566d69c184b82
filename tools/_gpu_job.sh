set -x
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2E_bench_n2.log 2> gpurun_out/r2E_bench_n2.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2E_bench_n2.log') if l.startswith('{')][-1])
    print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['stage_ms'], d['clocks'], d['gpu_launches'])
    for k,v in d['secondary'].items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['gpu_launches'])
except Exception as e: print("parse failed", e)
PY
tail -5 gpurun_out/r2E_bench_n2.err
