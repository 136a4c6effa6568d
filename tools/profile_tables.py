"""Turns the round's raw artefacts in gpurun_out/ into the tables committed under profiles/.
usage: python tools/profile_tables.py <tag> [round]   (after `gpurun -- bash tools/collect.sh <tag> ...`)"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
RND = sys.argv[2] if len(sys.argv) > 2 else "r2"


def display_name(kernel_name: str) -> str:
    """ncu kernel name -> the name bench.py's per-kernel timers use."""
    key = kernel_name.split("(")[0].replace("void ", "").replace("c3p::", "")
    if key.startswith("k_gather_mma"):
        args = key[key.index("<") + 1:key.index(">")].replace(" ", "").split(",")
        return "k_backward_input_tc" if args[1] in ("1", "true") else "k_forward_tc"
    if key.startswith("k_backward_filter2"):
        return "k_backward_filter_tc"
    return key.split("<")[0]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    hdr, units, rows = ncu_raw(os.path.join(G, f"prof_{TAG}.ncu-rep"))
    conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}

    def num(r, k):
        return float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] not in ("", "n/a") else 0.0

    traffic, lines = {}, []
    lines.append("| kernel | time ms | DRAM read MB | DRAM write MB | DRAM % | L2 hit % | tensor pipe % | L1/shared data pipe % "
                 "(LSU + tensor-core reads) | issue active % | warps active % | regs | busiest unit |")
    lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        key = display_name(kn)
        rd = num(r, "dram__bytes_read.sum") * conv[units[hdr.index("dram__bytes_read.sum")]]
        wr = num(r, "dram__bytes_write.sum") * conv[units[hdr.index("dram__bytes_write.sum")]]
        lsu = num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
        tcr = num(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
        units_pct = {
            "l1/shared-memory data pipe": lsu + tcr,
            "tensor pipe": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "instruction issue": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        }
        binder = max(units_pct, key=units_pct.get)
        if key in traffic:   # several launches of one kernel (k_group_items): keep the first
            key = key + "#" + str(sum(1 for k in traffic if k.startswith(key)))
        traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                        "time_ms": num(r, "gpu__time_duration.sum"),
                        "binder": f"{binder} ({units_pct[binder]:.0f} % busy under ncu)",
                        "pct": {k: round(v, 1) for k, v in units_pct.items()}}
        lines.append(f"| `{key}` | {num(r, 'gpu__time_duration.sum'):.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                     f"{units_pct['dram']:.1f} | {num(r, 'lts__t_sector_hit_rate.pct'):.1f} | {units_pct['tensor pipe']:.1f} | "
                     f"{lsu + tcr:.1f} ({lsu:.1f} + {tcr:.1f}) | {units_pct['instruction issue']:.1f} | "
                     f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{r[hdr.index('launch__registers_per_thread')]} | {binder} |")
    json.dump({"headline": traffic, "source": f"ncu --set full --clock-control none, gpurun_out/prof_{TAG}.ncu-rep, one launch each"},
              open(os.path.join(P, "kernel_traffic.json"), "w"), indent=1)
    open(os.path.join(P, f"{RND}_ncu_table.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    pairs = [(f"bench_{TAG}.json", f"{RND}_bench_headline.json"), (f"bench_{TAG}_reference.json", f"{RND}_bench_reference_arm.json"),
             (f"launches_{TAG}.csv", f"{RND}_launches_headline.csv"), (f"engine_timing_{TAG}.txt", f"{RND}_engine_timing.txt"),
             (f"sanitize_memcheck_{TAG}.log", f"{RND}_sanitize_memcheck.log"), (f"sanitize_racecheck_{TAG}.log", f"{RND}_sanitize_racecheck.log"),
             (f"pytest_{TAG}.log", f"{RND}_pytest_gpu.log")]
    for w in ("headline_b16", "s3dis_l1", "s3dis_l5", "modelnet_l2", "seg_net", "cls_net"):
        pairs.append((f"bench_{TAG}_{w}.json", f"{RND}_bench_{w}.json"))
    for src, dst in pairs:
        if os.path.exists(os.path.join(G, src)):
            shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
    bpath = os.path.join(G, f"bench_{TAG}.json")
    if os.path.exists(bpath):
        d = json.load(open(bpath))
        print("\nheadline:", d["value"], "points/s", d["ms_per_step"], "ms/step; e2e", d["e2e"]["value"], "; cpu",
              d.get("cpu_baseline", {}).get("value"))
        print("roofline:", {k: d["roofline"][k] for k in ("kernel", "bound", "frac_G", "frac_A", "frac_F")})
        for k, v in d["kernels"].items():
            print("  ", k, v)
        if "sweep" in d:
            json.dump(d["sweep"], open(os.path.join(P, f"{RND}_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
