"""Top stall sites of one kernel from an ncu report (SASS level, with dominant stall reasons).
usage: python tools/ncu_hotspots.py <report.ncu-rep> <kernel regex> [launch index] [top n]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1])
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ai, so, ie = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[si] or 0) for r in body)
    tot_by = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
    print("total samples", total, "instructions executed (warp)", sum(int(r[ie] or 0) for r in body))
    print({k: round(v / total, 3) for k, v in sorted(tot_by.items(), key=lambda kv: -kv[1])[:8]})
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
    for i in sorted(idx):
        r = body[i]
        reasons = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/total:5.1f}%  ex={r[ie]:>8s}  {r[so][:70]:70s} {reasons}")


main()
