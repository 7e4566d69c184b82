for rep in 1 2; do
for cfg in "0 1" "1 1" "0 2" "1 2"; do
  set -- $cfg
  LNASR_TC16_POLY=$1 LNASR_TC16_GROUPS=$2 timeout 300 python bench.py --no-cpu --no-secondary --steps 60 --warmup 10 > gpurun_out/r2G_$1_$2.log 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2G_$1_$2.log') if l.startswith('{')][-1])
print("poly $1 groups $2:", round(d['ms_per_step'],3), d['roofline']['stage_ms'], d['clocks']['sm_mhz'], round(d['e2e']['ms_per_step'],3))
PY
done; done
