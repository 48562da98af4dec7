#!/bin/bash
# Round-2 validation + A/B call (one B200, run under gpurun; everything lands in gpurun_out/ab_r2/).
#   1. the whole GPU suite on the default build
#   2. the G2-heavy parity tests again with the lane-pair G2 accumulation (SB_TUNE=11=4 and 11=3)
#   3. bench.py A/B of that kernel on the Groth16 2^20 workload (10 proofs each, no CPU arm)
#   4. bench.py --workload plonk at 2^18 on BLS12-381 with the three occupancy variants of the 12-limb G1 accumulation
cd "$(dirname "$0")/.."
O=gpurun_out/ab_r2; mkdir -p $O
date +%s > $O/t0
timeout 600 python -m pytest tests -q -m gpu --maxfail=5 --durations=12 > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_default.log | tee -a $O/summary.txt
G2K="ptau_goldens or vs_oracle or edge_cases or groth16_fused or synthetic_2_16 or bls12_381_synthetic or sharded_keys or plain_vs_table or witness_like"
for V in 4 3; do
  SB_TUNE=11=$V timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$G2K" --maxfail=5 > $O/pytest_pair$V.log 2>&1; echo "pytest pair$V rc=$?" | tee -a $O/summary.txt
  tail -1 $O/pytest_pair$V.log | tee -a $O/summary.txt
done
for T in "" "--tune 11=4" "--tune 11=3"; do
  N=$(echo "g16${T}" | tr -d ' =-'); timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $T > $O/$N.json 2> $O/$N.err; echo "$N rc=$?" >> $O/summary.txt
done
for T in "" "--tune 10=3" "--tune 10=2"; do
  N=$(echo "plonk18${T}" | tr -d ' =-'); timeout 240 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 --no-cpu-baseline $T > $O/$N.json 2> $O/$N.err; echo "$N rc=$?" >> $O/summary.txt
done
python - <<'PY' | tee -a gpurun_out/ab_r2/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_r2/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        b = d.get("breakdown_ms", {})
        print(f.split("/")[-1], "value %.2f e2e %.2f ms %.2f rint %.3f acc_g1 %.2f acc_g2 %s match %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_int"]["frac"], b.get("accumulate_g1", 0), b.get("accumulate_g2"), d.get("oracle_match")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "elapsed $(( $(date +%s) - $(cat $O/t0) )) s" | tee -a $O/summary.txt
