#!/bin/bash
# Round-2 ncu captures (run on the GPU box under gpurun; one GPU).  Writes CSV pages into gpurun_out/.
# 1. launch list of one serialised 2^20 Groth16 proof;  2. --set full of every kernel class of that proof;
# 3. --set full of the PLONK / fflonk round kernels at 2^14.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='regex:k_(accumulate|reduce|fold|fold_short|ntt_pass|qap_rows|digits|join_abc|window_sum|count_valid)'
SB_SERIAL=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python profiles/prof_one_proof.py > gpurun_out/launches_r2.log 2>&1
SB_SERIAL=1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k "$K" -c 80 -f -o gpurun_out/ncu_full_groth16_r2 python profiles/prof_one_proof.py > gpurun_out/ncu_full_groth16_r2.log 2>&1
ncu -i gpurun_out/ncu_full_groth16_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_groth16_r2.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
for P in plonk fflonk; do
  PROTO=$P LOGN=14 timeout 600 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_(pl|ff)_' -c 120 -f -o gpurun_out/ncu_full_${P}_r2 python profiles/prof_one_plonk.py > gpurun_out/ncu_full_${P}_r2.log 2>&1
  ncu -i gpurun_out/ncu_full_${P}_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_${P}_r2.csv 2>/dev/null
done
# keep the reports only if they fit the 64 MiB return budget
du -sm gpurun_out/*.ncu-rep
for f in gpurun_out/*.ncu-rep; do [ $(du -m "$f" | cut -f1) -gt 12 ] && rm -f "$f"; done
true
