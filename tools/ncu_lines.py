"""Where does a kernel stall?  Aggregates the sampling data of an ncu report by CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]

Prints the SASS size of the first matching kernel, its stall-reason totals, and the N source lines with the most samples
(needs a build with -lineinfo and a capture with --import-source on)."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kre}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, fname, func0 = None, "", None
    lines = {}
    totals = defaultdict(int)
    n_sass = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            if func0 is None:
                func0 = r[1]
            cur_ok = r[1] == func0
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not cur_ok or len(r) < len(hdr):
            continue
        d = dict(zip(hdr[4:], r[4:]))
        if r[0] == "":          # SASS row
            if r[2] not in ("...", "-"):
                n_sass += 1
            continue
        try:
            samples = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        key = f"{fname}:{r[0]}"
        rec = lines.setdefault(key, dict(samples=0, src=r[1], reasons=defaultdict(int)))
        rec["samples"] += samples
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    iv = int(v)
                except ValueError:
                    continue
                totals[k] += iv
                rec["reasons"][k] += iv
    print(func0)
    print(f"SASS instructions (sampled listing): {n_sass}")
    tot = sum(totals.values()) or 1
    print("stall totals:", ", ".join(f"{k[6:]}={v} ({100 * v / tot:.0f}%)" for k, v in sorted(totals.items(), key=lambda kv: -kv[1]) if v))
    for key, rec in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        reasons = ", ".join(f"{k[6:]}={v}" for k, v in sorted(rec["reasons"].items(), key=lambda kv: -kv[1]) if v)[:90]
        src = re.sub(r"\s+", " ", rec["src"]).strip()[:60]
        print(f"{rec['samples']:6d} {key:26s} {src:60s} {reasons}")


if __name__ == "__main__":
    main()
