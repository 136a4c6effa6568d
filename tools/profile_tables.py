"""Turns the round's raw artefacts in gpurun_out/ into the tables committed under profiles/.
usage: python tools/profile_tables.py [tag]  (after `gpurun -- bash tools/collect.sh <tag>`; tag defaults to r1)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
def display_name(kernel_name: str) -> str:
    """ncu kernel name -> the name bench.py's per-kernel timers use."""
    key = kernel_name.split("(")[0].replace("void ", "").replace("c3p::", "")
    if key.startswith("k_gather_mma"):
        args = key[key.index("<") + 1:key.index(">")].replace(" ", "").split(",")
        return "k_backward_input_tc" if args[1] in ("1", "true") else "k_forward_tc"
    if key.startswith("k_backward_filter2") or key.startswith("k_backward_filter_tc"):
        return "k_backward_filter_tc"
    return key


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    hdr, units, rows = ncu_raw(os.path.join(G, f"prof_{TAG}.ncu-rep"))
    g = lambda r, k: r[hdr.index(k)]
    conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    traffic, lines = {}, []
    lines.append("| kernel | time ms | DRAM read MB | DRAM write MB | DRAM % | L2 hit % | tensor pipe % | issue active % "
                 "| warps active % | regs |")
    lines.append("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        kn = g(r, "Kernel Name")
        key = display_name(kn)
        rd = float(g(r, "dram__bytes_read.sum").replace(",", "")) * conv[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(g(r, "dram__bytes_write.sum").replace(",", "")) * conv[units[hdr.index("dram__bytes_write.sum")]]
        if key in traffic:   # several launches of one kernel (k_group_items): keep the first
            key = key + "#" + str(sum(1 for k in traffic if k.startswith(key)))
        traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                        "time_ms": float(g(r, "gpu__time_duration.sum"))}
        lines.append(f"| `{key}` | {float(g(r, 'gpu__time_duration.sum')):.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                     f"{float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                     f"{float(g(r, 'lts__t_sector_hit_rate.pct')):.1f} | "
                     f"{float(g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
                     f"{float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
                     f"{float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
                     f"{g(r, 'launch__registers_per_thread')} |")
    json.dump({"headline": traffic}, open(os.path.join(P, "kernel_traffic.json"), "w"), indent=1)
    print("\n".join(lines))
    for src, dst in [(f"bench_{TAG}.json", "r1_bench_headline.json"), (f"bench_{TAG}_reference.json", "r1_bench_reference_arm.json"),
                     (f"launches_{TAG}.csv", "r1_launches_headline.csv"), (f"bench_{TAG}_s3dis_l1.json", "r1_bench_s3dis_l1.json"),
                     (f"bench_{TAG}_s3dis_l5.json", "r1_bench_s3dis_l5.json"), (f"bench_{TAG}_modelnet_l2.json", "r1_bench_modelnet_l2.json"),
                     (f"engine_timing_{TAG}.txt", "r1_engine_timing.txt")]:
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
    d = json.load(open(os.path.join(G, f"bench_{TAG}.json")))
    print("\nheadline:", d["value"], "points/s", d["ms_per_step"], "ms/step; e2e", d["e2e"]["value"], "; cpu", d["cpu_baseline"]["value"])
    print("roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac")})
    for k, v in d["kernels"].items():
        print("  ", k, v)
    for w in ("s3dis_l1", "s3dis_l5", "modelnet_l2"):
        f = os.path.join(G, f"bench_{TAG}_{w}.json")
        if os.path.exists(f):
            x = json.load(open(f))
            print(w, x["value"], x["ms_per_step"], "cpu", x.get("cpu_baseline", {}).get("value"))


if __name__ == "__main__":
    main()
