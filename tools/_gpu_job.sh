set -x
timeout 300 python tools/bench_em_large.py 200 4
timeout 300 python tools/bench_em_large.py 2000 2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2O_launches_em.csv python tools/bench_em_large.py 200 4 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2O_launches_em.csv | awk -F'","' 'NR>1{n=$5; sub(/\(.*/,"",n); t[n]+=$NF; c[n]++} END{for(k in t) printf "%-60s %4d launches  avg %9.1f us\n", k, c[k], t[k]/c[k]/1000}' | sort -k5 -nr | head -14
