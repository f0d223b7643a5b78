"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  prof_launches.csv     : `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of one bench step
  prof_gemm_fc1.ncu-rep : `ncu --set full` capture of the dominant GEMM launch (stage-2 fc1 + GELU)
"""
import csv, json, subprocess, sys, io, re
from collections import defaultdict
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def klass(name):
    if "tc_gemm_kernel" in name: return "gemm_tcgen05"
    if "tc_attn_kernel" in name: return "attn_tcgen05"
    if "tc_deform_kernel" in name: return "deform_tcgen05"
    if name.startswith("ln_") or "ln_bulk" in name or "ln_vec" in name: return "layernorm"
    return "glue"

rows = []
with open(ROOT / "gpurun_out" / "prof_launches.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"^void ", "", r["Kernel Name"]).split("(")[0]
    rows.append((int(r["ID"]), name, r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
with open(OUT / f"{tag}_launches.csv", "w") as f:
    f.write("id,kernel,grid,block,us\n")
    for r in rows: f.write(f'{r[0]},"{r[1]}","{r[2]}","{r[3]}",{r[4]:.2f}\n')
agg = defaultdict(lambda: [0, 0.0]); per = defaultdict(lambda: [0, 0.0])
for _, name, _, _, us in rows:
    a = agg[klass(name)]; a[0] += 1; a[1] += us
    b = per[name]; b[0] += 1; b[1] += us
tot = sum(a[1] for a in agg.values())
bench = None
bl = ROOT / "gpurun_out" / "prof_plain.log"
if bl.exists():
    js = [l for l in bl.read_text().splitlines() if l.startswith("{")]
    if js: bench = json.loads(js[-1])
md = [f"# {tag}: ncu launch list of one bench step (batch 16, 1024^2, fp16 tensor-core path, direct launches)\n",
      "Command: `BRN_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1746 -c 291 --csv python bench.py "
      "--steps 1 --warmup 3 --no-cpu-baseline --no-latency` (after the same command exited 0 without ncu).  Per-launch times under "
      "ncu are cold-cache and serialised: compare the SHARES with the live CUDA-event shares of `bench.py` (right column), not the absolutes.\n",
      f"{len(rows)} launches, {tot/1e3:.2f} ms summed.\n", "| class | launches | ms (ncu) | share (ncu) | share (bench.py CUDA events) |", "|---|---|---|---|---|"]
live = bench["roofline"]["classes_ms"] if bench else {}
ltot = sum(live.values()) if live else 0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lv = f"{100*live.get(k,0)/ltot:.1f} %" if ltot else "n/a"
    md.append(f"| {k} | {a[0]} | {a[1]/1e3:.2f} | {100*a[1]/tot:.1f} % | {lv} |")
md += ["", "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, a in sorted(per.items(), key=lambda kv: -kv[1][1])[:14]:
    md.append(f"| `{k}` | {a[0]} | {a[1]/1e3:.2f} | {100*a[1]/tot:.1f} % |")
# full capture of the dominant GEMM
rep = ROOT / "gpurun_out" / "prof_gemm_fc1.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, u, v = rr[0], rr[1], rr[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    vals = {}
    with open(OUT / f"{tag}_ncu_full_tc_gemm_fc1.csv", "w") as f:
        f.write("metric,unit,value\n")
        for i, k in enumerate(h):
            if k in want or k == "Kernel Name":
                f.write(f'"{k}","{u[i]}","{v[i]}"\n'); vals[k] = (u[i], v[i])
    def tobytes(k):
        unit, val = vals[k]; val = float(val)
        return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    dram = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    M, N, K = 81920, 3072, 768
    alg = M * K * 2 + M * N * 2 + N * K * 2
    json.dump({"kernel": "tc_gemm_kernel<2, 3> (2-CTA cluster, GELU16 epilogue), stage-2 fc1 + erf-GELU of the merged backbone pass (M=81920 N=3072 K=768, fp16 in/out)",
               "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": alg,
               "note": f"ncu --set full capture {tag}: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM "
                       f"(stage-2 fc1, 18 launches per step); algorithmic bytes of that launch = {alg/1e6:.0f} MB (A + out + W); part of the "
                       "output is still dirty in L2 when the kernel ends",
               "source": f"profiles/{tag}_ncu_full_tc_gemm_fc1.csv"}, open(OUT / "traffic.json", "w"), indent=1)
    md += ["", f"## `ncu --set full` of the dominant GEMM launch (stage-2 fc1 + GELU, M={M} N={N} K={K})", "",
           f"DRAM traffic {dram/1e6:.0f} MB vs algorithmic {alg/1e6:.0f} MB (no re-reads); metrics in `{tag}_ncu_full_tc_gemm_fc1.csv`:", ""]
    for k, (uu, vv) in vals.items():
        if k != "Kernel Name": md.append(f"* `{k}` = {vv} {uu}")
(OUT / f"{tag}_summary.md").write_text("\n".join(md) + "\n")
print("\n".join(md))
