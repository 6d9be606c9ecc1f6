"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/: key metrics, stall reasons, top SASS lines.

    python tools/ncu_summary.py gpurun_out/prof_embed_config2_r01.ncu-rep config2 r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

rep, workload, rnd = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("raw")
hdr, units, rows = raw[0], raw[1], raw[2:]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
lines = [f"# ncu --set full summary: {workload}, round {rnd}", "",
         f"source: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on, one launch; "
         "replayed ~40x, cold caches: use for ratios and traffic, not for the bench number)", ""]
summary = {}
for r in rows:
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {r[i]} | {units[i]} |")
            summary[k] = r[i]
    lines.append("")
    lines.append("stall reasons (warp cycles per issued instruction):")
    lines.append("")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    for v, n in sorted(st, reverse=True)[:8]:
        lines.append(f"* {n}: {v:.2f}")
    lines.append("")

    def f(k):
        i = hdr.index(k)
        v = float(r[i])
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
    json.dump({"workload": workload, "round": rnd, "kernel": r[hdr.index("Kernel Name")],
               "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
               "dram_bytes_per_launch": traffic, "gpu_time_us_under_ncu": summary.get("gpu__time_duration.sum")},
              open(os.path.join(ROOT, "profiles", f"traffic_{workload}.json"), "w"), indent=1)
    lines.append(f"DRAM traffic per launch: {traffic / 1e6:.1f} MB (read {f('dram__bytes_read.sum') / 1e6:.1f} + write {f('dram__bytes_write.sum') / 1e6:.1f})")
    lines.append("")

src = page("source")
kern, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = []
        kern.append(cur)
    elif r and r[0] != "Address" and cur is not None and len(r) > 6:
        cur.append(r)
if kern:
    k = kern[0]
    tot = sum(int(x[5]) for x in k)
    byop = collections.Counter()
    for x in k:
        parts = x[1].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        byop[op.split(".")[0]] += int(x[5])
    lines.append(f"SASS: {len(k)} instructions, {tot} warp-instructions executed; by opcode:")
    lines.append("")
    lines.append(", ".join(f"{op} {100 * c / tot:.1f}%" for op, c in byop.most_common(14)))
    lines.append("")
    mem = [x[1].split()[0] if not x[1].strip().startswith("@") else x[1].split()[1] for x in k]
    proof = sorted({m for m in mem if m.startswith(("UBLKCP", "LDG", "STG", "SYNCS", "LDS"))})
    lines.append("memory / async-copy mnemonics present: " + ", ".join(proof))
    lines.append("")
    lines.append("top stall-sample instructions:")
    lines.append("")
    lines.append("```")
    for x in sorted(k, key=lambda x: -int(x[2]))[:12]:
        lines.append(f"{x[2]:>6} samples  {x[5]:>9} exec  {x[1].strip()[:100]}")
    lines.append("```")
out = os.path.join(ROOT, "profiles", f"ncu_embed_{workload}_{rnd}.md")
open(out, "w").write("\n".join(lines) + "\n")
print(open(out).read())
