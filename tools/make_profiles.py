#!/usr/bin/env python
"""Turns the .ncu-rep files of tools/profile_job.sh into the committed evidence: profiles/<round>_ncu_summary.md (one row per kernel),
profiles/<round>_traffic.json (DRAM bytes per launch, read back by bench.py as roofline.traffic) and a copy of the launch list.
   python tools/make_profiles.py r02"""
import csv, io, json, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum", "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
     "l2req": "lts__t_requests_srcunit_tex.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size"}
try:
    PEAK_GBS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    PEAK_GBS = 6650.0
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, m in M.items():
            if m in hdr:
                v = float(r[hdr.index(m)].replace(",", "")) if r[hdr.index(m)] not in ("", "n/a") else None
                u = units[hdr.index(m)]
                if v is not None and u in UNIT and k in ("dur", "rd", "wr"):
                    v *= UNIT[u]
                d[k] = v
        res.append(d)
    return res


def short(k):
    k = k.replace("npair::", "")
    return k[:k.index("(")] if "(" in k else k


allrows = []
for tag in ("main", "lsel", "gsel"):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{R}_{tag}.ncu-rep")
    if os.path.exists(rep):
        for d in rows_of(rep):
            d["capture"] = tag
            allrows.append(d)
lines = [f"# ncu --set full captures, round {R[1:]} (B = 8192, D = 512, fp16x2; `tools/profile_job.sh`, `--clock-control none`)", "",
         "Durations under ncu are cold-cache and serialised; the live per-phase CUDA-event times are in the bench lines.", "",
         "| capture | kernel | grid x block | regs | duration µs | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of the measured copy peak | tensor pipe active % | issue active % | warps active % | warp instr (M) | L2 requests (M) |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for d in allrows:
    gbs = (d["rd"] + d["wr"]) / (d["dur"] * 1e-6) / 1e9 if d.get("dur") else 0
    lines.append(f"| {d['capture']} | `{short(d['kernel'])}` | {int(d['grid'])} x {int(d['block'])} | {int(d['regs'])} | {d['dur']:.1f} | {d['rd'] / 1e6:.1f} | {d['wr'] / 1e6:.1f} | "
                 f"{gbs:.0f} | {100.0 * gbs / PEAK_GBS:.1f} | {d.get('tensor') or 0:.1f} | {d.get('issue') or 0:.1f} | {d.get('warps') or 0:.1f} | {(d.get('inst') or 0) / 1e6:.1f} | {(d.get('l2req') or 0) / 1e6:.2f} |")
    traffic.setdefault(d["capture"], []).append({"kernel": short(d["kernel"]), "dram_bytes": d["rd"] + d["wr"], "dram_read": d["rd"], "dram_write": d["wr"], "duration_us": d["dur"]})
open(os.path.join(ROOT, "profiles", f"{R}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
kern = {}
for t in traffic.get("main", []):
    n = t["kernel"]
    key = ("sim_gemm" if "split_gemm" in n else "grad_gemm" if "fused_grad" in n else "row_pass" if "lse_rows" in n else "prep_reduce" if "prep_reduce" in n
           else "split" if "split_kernel" in n else "thresholds" if "thresholds" in n else n)
    kern[key] = t
json.dump({"workload": {"B": 8192, "D": 512, "precision": "fp16x2", "world": 1, "mining": "usage"}, "source": f"gpurun_out/prof_{R}_main.ncu-rep (ncu --set full --clock-control none)",
           "kernels": kern, "selects": {"local_relative": traffic.get("lsel", []), "global_relative": traffic.get("gsel", [])}},
          open(os.path.join(ROOT, "profiles", f"{R}_traffic.json"), "w"), indent=1)
src = os.path.join(ROOT, "gpurun_out", f"launches_{R}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(ROOT, "profiles", f"{R}_launches.csv"))
print("\n".join(lines))
