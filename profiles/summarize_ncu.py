"""Turn the ncu captures of a round (gpurun_out/*.ncu-rep, launch-list CSV) into the committed summaries.
    python profiles/summarize_ncu.py r02
Reads with `ncu -i ... --page raw --csv` (works without a GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, row)) for row in rows[2:]], dict(zip(hdr, units))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lines, traffic = [], {}
    for f in sorted(os.listdir(OUT)):
        if not (f.startswith(tag + "_prof_") and f.endswith(".ncu-rep")):
            continue
        recs, units = raw(os.path.join(OUT, f))
        for rec in recs:
            name = rec.get("Kernel Name", "?").split("(")[0]
            lines.append("%s  [%s]  grid %s block %s" % (name, f, rec.get("Grid Size"), rec.get("Block Size")))
            for key, label in WANT:
                if key in rec and rec[key] != "":
                    lines.append("    %-26s %s %s" % (label, rec[key], units.get(key, "")))
            try:
                def tobytes(k):
                    v, u = float(rec[k].replace(",", "")), units.get(k, "byte")
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
                traffic.setdefault(name.split("<")[0].split("::")[-1], []).append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
            except Exception:
                pass
            lines.append("")
    open(os.path.join(ROOT, "profiles", tag + "_ncu_full_summary.txt"), "w").write("\n".join(lines))
    print("\n".join(lines))
    # launch list
    for f in sorted(os.listdir(OUT)):
        if not (f.startswith(tag + "_launches_") and f.endswith(".csv")):
            continue
        rows = list(csv.reader(open(os.path.join(OUT, f))))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= iv:
                continue
            name = r[ik].split("(")[0]
            v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1
            agg[name][1] += v
        tot = sum(v[1] for v in agg.values())
        out = ["# %s: every kernel launch of the command, device time from ncu --metrics gpu__time_duration.sum (cold-cache, serialised: compare shares)" % f,
               "# total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values()))]
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.append("%-70s n=%5d  %10.3f ms  %5.1f %%" % (k[-70:], v[0], v[1], 100 * v[1] / tot))
        open(os.path.join(ROOT, "profiles", f.replace(".csv", "_summary.txt")), "w").write("\n".join(out) + "\n")
        print("\n".join(out[:16]))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
