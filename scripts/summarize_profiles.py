"""Turn the scratch ncu output in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>      e.g.  r01_v2

  gpurun_out/launches.csv        -> profiles/<tag>_launches.md   (per-kernel launches / total time / share)
  gpurun_out/prof_*.ncu-rep      -> profiles/<tag>_ncu_full.json (selected raw metrics per captured launch)
  largest captured GEMM launch   -> profiles/gemm_traffic.json   (DRAM traffic per launch, read by bench.py)
"""
import collections
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("DRIN_PROFILE_OUT", os.path.join(ROOT, "profiles"))     # on the GPU box: a directory under gpurun_out/

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_uniform.sum",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").strip()[:100]


def launches(tag, note):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    text = open(path).read()
    text = text[text.index('"ID"'):]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(text)):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(row["Kernel Name"])
        ns = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("drin::"))
    lines = [f"# ncu launch list, {tag}", "", note, "",
             f"Total {total / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches; drin:: kernels {ours / total * 100:.1f}% "
             "(the rest is synthetic-data generation by torch). Times are cold-cache/serialised: compare shares.", "",
             "| kernel | launches | total us | share | share of drin:: |", "|---|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        so = f"{ns / ours * 100:.1f}%" if k.startswith("drin::") else ""
        lines.append(f"| {k} | {n} | {ns / 1e3:.1f} | {ns / total * 100:.1f}% | {so} |")
    with open(os.path.join(PROF, f"{tag}_launches.md"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as fh:
        fh.write(text)


def ncu_raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if r.returncode != 0:
        return []
    text = r.stdout
    text = text[text.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(text)))
    head, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        rec = {"kernel": short(d[head.index("Kernel Name")])}
        for m in METRICS:
            if m in head:
                i = head.index(m)
                rec[m] = f"{d[i]} {units[i]}".strip()
        out.append(rec)
    return out


def to_bytes(s):
    v, u = s.split()[:2]
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    os.makedirs(PROF, exist_ok=True)
    launches(tag, note)
    full = {}
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        recs = ncu_raw(rep)
        if recs:
            full[os.path.basename(rep)[:-8]] = recs
    if full:
        with open(os.path.join(PROF, f"{tag}_ncu_full.json"), "w") as fh:
            json.dump(full, fh, indent=1)
    gem = [r for r in full.get("prof_gemm", []) if "dram__bytes_read.sum" in r]
    if gem:
        big = max(gem, key=lambda r: to_bytes(r["dram__bytes_read.sum"]) + to_bytes(r["dram__bytes_write.sum"]))
        rd, wr = to_bytes(big["dram__bytes_read.sum"]), to_bytes(big["dram__bytes_write.sum"])
        with open(os.path.join(PROF, "gemm_traffic.json"), "w") as fh:
            json.dump({
                "source": f"gpurun_out/prof_gemm.ncu-rep (ncu --set full --clock-control none), {tag}",
                "kernel": big["kernel"] + " entity-image projection: M=45056 (4096 mentions x 11 candidates), N=768, K=2048, split-bf16 planes",
                "traffic_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                "algorithmic_bytes_per_launch": 45056 * 2048 * 4 + 768 * 2048 * 4 + 45056 * 768 * 4,
                "note": "algorithmic = A planes 45056x2048x(2+2) B + W planes 768x2048x4 B + fp32 output 45056x768x4 B",
                "tensor_pipe_active_pct": big.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "duration_under_ncu": big.get("gpu__time_duration.sum"),
            }, fh, indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
