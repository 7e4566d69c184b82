# Condense gpurun_out/fin_* (tools/gpu_round2_final.sh) into profiles/r2_* on the authoring side.
set -e
G=gpurun_out
python - <<'PY'
import json
t=json.load(open('gpurun_out/fin_ncu_traffic.json'))
for w in t:
    t[w]['source']=t[w]['source'].replace('profiles/fin_%s_ncu_full.txt'%w, 'profiles/r2_%s_ncu_full.txt (ncu --set full of bench.py --workload %s --steps 1, end of round 2; tools/gpu_round2_final.sh)'%(w,w))
json.dump(t,open('profiles/ncu_traffic.json','w'),indent=1)
for src,dst in (('fin_bench_c4.json','r2_bench_c4_n1.json'),('fin_bench_ref.json','r2_bench_reference_arm_c4.json')):
    d=json.loads([l for l in open('gpurun_out/'+src) if l.startswith('{')][-1]); json.dump(d,open('profiles/'+dst,'w'))
PY
for w in c2 c4; do
  sed "s#gpurun_out/fin_$w.ncu-rep#end of round 2: ncu --set full -k <dominant kernels> -s 9 -c 3 of bench.py --workload $w --steps 1 --warmup 3 (tools/gpu_round2_final.sh)#" $G/fin_${w}_ncu_full.txt > profiles/r2_${w}_ncu_full.txt
  cp $G/fin_launches_$w.csv profiles/r2_launches_$w.csv
done
cp $G/fin_launches_em.csv profiles/r2_launches_em_200x4.csv
cp $G/fin_mfcc_sweep.md profiles/r2_c5_mfcc_sweep.md
python tools/ncu_summary.py $G/fin_mfcc.ncu-rep | sed 's#gpurun_out/fin_mfcc.ncu-rep#end of round 2: mfcc_frames_kernel at 1 s x 4096 (tools/mfcc_sweep.py --only 1x4096)#' > profiles/r2_mfcc_ncu_full.txt
python tools/ncu_hotspots.py $G/fin_mfcc.ncu-rep mfcc_frames 45 | sed 's#gpurun_out/fin_mfcc.ncu-rep#end of round 2, 1 s x 4096#' > profiles/r2_mfcc_source_hotspots.txt
python tools/ncu_summary.py $G/fin_mfcc_b1.ncu-rep | sed 's#gpurun_out/fin_mfcc_b1.ncu-rep#end of round 2: one utterance of 1 s (tools/mfcc_sweep.py --only 1x1), the three kernels of one call#' > profiles/r2_c5_batch1_ncu_full.txt
python tools/ncu_summary.py $G/fin_emm.ncu-rep | sed 's#gpurun_out/fin_emm.ncu-rep#end of round 2: em_moments_mma_kernel, 200 states x 4 mixtures, 300 333 frames (tools/bench_em_large.py 200 4)#' > profiles/r2_em_moments_ncu_full.txt
python tools/ncu_hotspots.py $G/fin_emm.ncu-rep em_moments_mma 30 | sed 's#gpurun_out/fin_emm.ncu-rep#end of round 2, 200 x 4#' > profiles/r2_em_moments_source_hotspots.txt
