"""Per-kernel summary of an `ncu --set full --page raw --csv` export (one row per profiled launch).

    python profiles/summarize_ncu.py raw.csv [out.csv] [traffic.json]

Writes one line per distinct kernel (mean over its launches): duration, registers, grid x block, achieved occupancy,
integer/FMA pipe utilisation (IMAD.WIDE issues on the "fmaheavy" pipe), issue slots, executed warp instructions, DRAM
bytes read + written per launch and the HBM GB/s they amount to, L2 hit rate, and the dominant stall reasons.  With a
third argument the DRAM bytes per launch are also written as JSON (bench.py reads them for `roofline.traffic`)."""
import collections
import csv
import json
import re
import sys

WANT = [
    ("ms", "gpu__time_duration.sum"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("pipe_fmaheavy_pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("pipe_fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("inst_fmaheavy", "sm__inst_executed_pipe_fmaheavy.sum"),
    ("inst_fma", "sm__inst_executed_pipe_fma.sum"),
    ("dram_read_B", "dram__bytes_read.sum"),
    ("dram_write_B", "dram__bytes_write.sum"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("stall_wait", "smsp__pcsamp_warps_issue_stalled_wait"),
    ("stall_math_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"),
    ("stall_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard"),
    ("stall_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard"),
    ("stall_barrier", "smsp__pcsamp_warps_issue_stalled_barrier"),
    ("stall_not_selected", "smsp__pcsamp_warps_issue_stalled_not_selected"),
    ("stall_selected", "smsp__pcsamp_warps_issue_stalled_selected"),
]
SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"\bvoid |sb::|cub::CUB_\w+::|detail::|radix::", "", name)
    return name[:72]


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col.get("Kernel Name")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) <= kn:
            continue
        a = agg.setdefault(short(r[kn]), collections.defaultdict(list))
        for key, metric in WANT:
            i = col.get(metric)
            if i is None or i >= len(r):
                continue
            v = num(r[i])
            if v is None:
                continue
            a[key].append(v * SCALE.get(units[i], 1.0) if key in ("ms", "dram_read_B", "dram_write_B") else v)
    keys = [k for k, _ in WANT]
    out = [["kernel", "launches"] + keys + ["dram_B_per_launch", "hbm_GBps"]]
    traffic = {}
    for name, a in agg.items():
        n = len(a["ms"]) or 1
        mean = {k: (sum(a[k]) / len(a[k]) if a[k] else None) for k in keys}
        stalls = {k: mean[k] for k in keys if k.startswith("stall_") and mean[k]}
        tot = sum(stalls.values()) or 1.0
        for k in stalls:
            mean[k] = 100.0 * stalls[k] / tot                       # share of the listed stall samples, %
        dram = (mean["dram_read_B"] or 0) + (mean["dram_write_B"] or 0)
        gbps = dram / (mean["ms"] * 1e-3) / 1e9 if mean["ms"] else None
        out.append([name, n] + [("%.6g" % mean[k] if mean[k] is not None else "") for k in keys] + ["%.6g" % dram, ("%.5g" % gbps if gbps else "")])
        traffic[name] = {"dram_bytes_per_launch": dram, "ms": mean["ms"], "launches": n}
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
