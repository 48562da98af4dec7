#!/bin/bash
# Round-2 closing call (one B200, under gpurun; output in gpurun_out/final_r2/): the GPU suite on the final defaults, the
# bench lines of record (Groth16 2^20 with the CPU arm, PLONK BLS12-381 2^20), an A/B of the BN254 G1 accumulation at 3 and
# 2 CTAs/SM, and one ncu --set full capture of the BLS12-381 G1 accumulation at its new default (2 CTAs/SM).
cd "$(dirname "$0")/.."
O=gpurun_out/final_r2; mkdir -p $O
date +%s > $O/t0
timeout 400 python -m pytest tests -q -m gpu --maxfail=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -1 $O/pytest.log | tee -a $O/summary.txt
timeout 200 python bench.py > $O/bench_groth16.json 2> $O/bench_groth16.err; echo "bench groth16 rc=$?" >> $O/summary.txt
timeout 200 python bench.py --workload plonk --log-n 20 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_plonk20.json 2> $O/bench_plonk20.err; echo "bench plonk20 rc=$?" >> $O/summary.txt
for V in 3 2; do
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tune 12=$V > $O/g16_bn_g1_minb$V.json 2> $O/g16_bn_g1_minb$V.err; echo "g16 12=$V rc=$?" >> $O/summary.txt
done
PROTO=plonk LOGN=18 timeout 200 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_accumulate' -c 3 -f -o $O/ncu_bls_g1_acc python profiles/prof_one_plonk.py > $O/ncu_bls_g1_acc.log 2>&1
ncu -i $O/ncu_bls_g1_acc.ncu-rep --page raw --csv > $O/ncu_bls_g1_acc.raw.csv 2>/dev/null
python profiles/summarize_ncu.py $O/ncu_bls_g1_acc.raw.csv $O/ncu_bls_g1_acc.summary.csv $O/ncu_bls_g1_acc.traffic.json >> $O/summary.txt 2>&1
rm -f $O/*.ncu-rep
python - <<'PY' | tee -a gpurun_out/final_r2/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/final_r2/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "value" not in d: continue
        b = d.get("breakdown_ms", {})
        print(f.split("/")[-1], "value %.2f e2e %.2f ms %.2f rint %.3f acc_g1 %.2f acc_g2 %s match %s live %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_int"]["frac"], b.get("accumulate_g1", 0), b.get("accumulate_g2"), d.get("oracle_match"), d.get("oracle_live_match"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "elapsed $(( $(date +%s) - $(cat $O/t0) )) s" | tee -a $O/summary.txt
