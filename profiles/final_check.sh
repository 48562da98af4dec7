timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_final.json
python - <<'PY'
import json
r=json.load(open("gpurun_out/bench_r1_reference.json")); print("REF", r["value"], r["cpu_baseline"]["sample"][:120])
d=json.load(open("gpurun_out/bench_r1_final.json")); print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_int"]["frac"], d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"])
PY
