#!/bin/bash
# Round-2 ncu captures (run on the GPU box under gpurun; one GPU).  Writes raw CSV pages + per-kernel summaries into gpurun_out/.
#   1. launch list of one serialised 2^20 Groth16 proof (per-launch durations)
#   2. --set full of every kernel class of that proof
#   3. --set full of the PLONK (BLS12-381) and fflonk (BN254) proofs at 2^LOGN_PL (round kernels + their MSM / NTT kernels)
#   4. --set full of the 2^24 NTT round trip (BASELINE config #3)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOGN_PL=${LOGN_PL:-18}
K='regex:k_(accumulate|fold_short|fold|axis_sum|axis_tree|ws_chunks|ws_final|ntt_pass|qap_rows|digits|join_abc|count_valid)'
SB_SERIAL=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python profiles/prof_one_proof.py > gpurun_out/launches_r2.log 2>&1
SB_SERIAL=1 timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "$K" -c 90 -f -o gpurun_out/ncu_full_groth16_r2 python profiles/prof_one_proof.py > gpurun_out/ncu_full_groth16_r2.log 2>&1
ncu -i gpurun_out/ncu_full_groth16_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_groth16_r2.raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/ncu_full_groth16_r2.raw.csv gpurun_out/ncu_full_groth16_r2.summary.csv gpurun_out/ncu_traffic_groth16_r2.json
for P in plonk fflonk; do
  PROTO=$P LOGN=$LOGN_PL timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_(pl|ff|accumulate|axis_sum|ntt_pass|fold_short)' -c 160 -f -o gpurun_out/ncu_full_${P}_r2 python profiles/prof_one_plonk.py > gpurun_out/ncu_full_${P}_r2.log 2>&1
  ncu -i gpurun_out/ncu_full_${P}_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_${P}_r2.raw.csv 2>/dev/null
  python profiles/summarize_ncu.py gpurun_out/ncu_full_${P}_r2.raw.csv gpurun_out/ncu_full_${P}_r2.summary.csv gpurun_out/ncu_traffic_${P}_r2.json
done
NTT_LOGN=24 timeout 600 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_ntt_pass' -c 6 -f -o gpurun_out/ncu_full_ntt24_r2 python profiles/prof_one_ntt.py > gpurun_out/ncu_full_ntt24_r2.log 2>&1
ncu -i gpurun_out/ncu_full_ntt24_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_ntt24_r2.raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/ncu_full_ntt24_r2.raw.csv gpurun_out/ncu_full_ntt24_r2.summary.csv
# keep the reports only if they fit the 64 MiB return budget
du -sm gpurun_out/*.ncu-rep
for f in gpurun_out/*.ncu-rep; do [ $(du -m "$f" | cut -f1) -gt 10 ] && rm -f "$f"; done
true
