"""Per-source-line view of one `ncu --set full --import-source on` capture (kernels built with -lineinfo).

    python profiles/ncu_lines.py gpurun_out/<capture>.ncu-rep [top]

Aggregates the SASS rows of `ncu --page source --print-source cuda,sass --csv` by CUDA source line: stall samples,
warp instructions executed, excessive shared-memory wavefronts (bank conflicts) and the dominant stall reasons.
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hdr]
    col = {name: i for i, name in reversed(list(enumerate(h)))}
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = defaultdict(lambda: {"samples": 0, "inst": 0, "exc": 0, "stalls": defaultdict(int), "src": ""})
    line, src = None, ""
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        if r[0]:
            line, src = r[0], r[1]
        if not r[2]:
            continue
        a = agg[line]
        a["src"] = src
        try:
            a["samples"] += int(r[col["# Samples"]])
            a["inst"] += int(r[col["Instructions Executed"]])
            a["exc"] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
            for i, n in stall_cols:
                a["stalls"][n] += int(r[i] or 0)
        except ValueError:
            pass
    tot = sum(a["samples"] for a in agg.values()) or 1
    toti = sum(a["inst"] for a in agg.values()) or 1
    print("%s: %d samples, %d warp instructions" % (rep, tot, toti))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("L%-5s %5.1f%% smp %5.1f%% inst  bankx %-10d %-44s | %s"
              % (line, 100.0 * a["samples"] / tot, 100.0 * a["inst"] / toti, a["exc"],
                 " ".join("%s=%d%%" % (n[6:], 100 * v // max(1, a["samples"])) for n, v in st), a["src"].strip()[:90]))


if __name__ == "__main__":
    main()
