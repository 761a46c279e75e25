"""Turn the raw ncu outputs in gpurun_out/ (tools/gpu_r2_ncu*.sh) into the committed summaries under profiles/.

    python tools/summarize_profiles.py r2        # build container; needs ncu for the .ncu-rep files

Inputs (per tag):  launches_<tag>.csv        ncu launch list of `bench.py --steps 2 --warmup 1` (our kernels, 2 timed steps)
                   forward_dram_<tag>.csv    dram__bytes_{read,write}.sum + duration of every launch of one forward
                   gcfn_<tag>_final.ncu-rep  `--set full` capture of two k_gcfn launches
                   attn_<tag>.ncu-rep        `--set full` capture of one k_attn_relpos launch
"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
B, Tp, F = 32, 8000, 128


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    return h, [r for r in rows[hi + 1:] if len(r) == len(h)]


def short(name):
    n = name.replace("void ", "").replace("sepref::", "")
    n = re.sub(r"\((?:CUtensorMap|const|float|sepref|simt|tc|double|int).*", "", n)
    n = re.sub(r"TokCfg<([^>]*)>", lambda m: "TokCfg<" + m.group(1).replace(" ", "").replace("(int)", "").replace("(bool)", "") + ">", n)
    return n.replace("(int)", "").replace("(bool)", "")[:100]


def to_us(v, unit):
    unit = unit.lower()
    return v / 1000 if unit.startswith("n") else v if unit.startswith("u") else v * 1000 if unit.startswith("m") else v * 1e6


def to_bytes(v, unit):
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit.lower()]


# ---- launch list of the bench command -------------------------------------------------------------------------------
h, body = read_ncu_csv(os.path.join(G, f"launches_{tag}.csv"))
ci = {n: i for i, n in enumerate(h)}
acc = collections.OrderedDict()
tot = 0.0
for r in body:
    us = to_us(float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]])
    a = acc.setdefault(short(r[ci["Kernel Name"]]), [0, 0.0])
    a[0] += 1; a[1] += us; tot += us
nl = sum(a[0] for a in acc.values())
with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag} - ncu launch list of `bench.py --steps 2 --warmup 1` (the 2 timed forwards, B=32, Base, FP16 operands)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 500 -c 500` on `bench.py --steps 2 --warmup 1 --no-cuda-graph` (skips the model-level call that\n"
            "prepares the inputs and the warm-up step).  Per-launch times are cold-cache and serialised: compare SHARES with the\n"
            "CUDA-event numbers of the bench line (`kernel_ms`), not absolutes.  Every launch is one of this repo's kernels - no\n"
            "cuBLAS / cuDNN / Triton kernel runs inside a step.\n\n"
            f"{nl} launches, {tot/1000:.2f} ms summed ({tot/2000:.2f} ms per forward).\n\n"
            "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us:.0f} | {100*us/tot:.1f} % |\n")
subprocess.run(["cp", os.path.join(G, f"launches_{tag}.csv"), os.path.join(OUT, f"{tag}_launches_raw.csv")], check=True)

# ---- DRAM traffic of one forward, per kernel -------------------------------------------------------------------------
h, body = read_ncu_csv(os.path.join(G, f"forward_dram_{tag}.csv"))
ci = {n: i for i, n in enumerate(h)}
per = collections.OrderedDict()
for r in body:
    d = per.setdefault(r[ci["ID"]], {"name": short(r[ci["Kernel Name"]])})
    v = float(r[ci["Metric Value"]].replace(",", ""))
    m = r[ci["Metric Name"]]
    d[m] = to_bytes(v, r[ci["Metric Unit"]]) if "bytes" in m else to_us(v, r[ci["Metric Unit"]])
kern = collections.OrderedDict()
for d in per.values():
    k = kern.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    k[0] += 1; k[1] += d["dram__bytes_read.sum"]; k[2] += d["dram__bytes_write.sum"]; k[3] += d["gpu__time_duration.sum"]
rd = sum(k[1] for k in kern.values()); wr = sum(k[2] for k in kern.values()); us = sum(k[3] for k in kern.values())
alg_total = 83.0 * 2 * 4 * F * B * Tp                         # one read + one write of [tok, F] per fused block
alg_gcfn = 2 * 4 * F * 41.5 * B * Tp
g = [k for n, k in kern.items() if "k_gcfn" in n]
grd, gwr, gn = sum(k[1] for k in g), sum(k[2] for k in g), sum(k[0] for k in g)
json.dump({"launches": gn, "dram_bytes_per_launch": (grd + gwr) / max(gn, 1), "dram_read_bytes_per_forward": grd, "dram_write_bytes_per_forward": gwr,
           "algorithmic_bytes_per_forward": alg_gcfn, "algorithmic_bytes_per_launch": alg_gcfn / max(gn, 1),
           "whole_forward": {"launches": len(per), "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes": alg_total,
                             "ncu_time_us": us}},
          open(os.path.join(OUT, f"{tag}_gcfn_traffic.json"), "w"), indent=1)
with open(os.path.join(OUT, f"{tag}_forward_traffic.md"), "w") as f:
    f.write(f"# {tag} - DRAM traffic of one separator forward (B=32, Base, FP16 operands), per kernel\n\n"
            "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_ -s 246 -c 246\n"
            "python tools/one_forward.py` (second forward; per-stage outputs off: 246 launches).\n\n"
            f"Whole forward: read {rd/1e9:.2f} GB + write {wr/1e9:.2f} GB = **{(rd+wr)/1e9:.2f} GB** against {alg_total/1e9:.2f} GB algorithmic "
            f"(SURVEY 8d: 83 block passes x read + write of [frame, F] fp32): ratio {(rd+wr)/alg_total:.2f}.  "
            f"GCFN launches alone: {(grd+gwr)/1e9:.2f} GB against {alg_gcfn/1e9:.2f} GB (ratio {(grd+gwr)/alg_gcfn:.2f}; reads that hit in the 126 MB L2 do not reach DRAM).\n\n"
            "| kernel | launches | DRAM read MB | DRAM write MB | ncu us |\n|---|---:|---:|---:|---:|\n")
    for n, k in sorted(kern.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        f.write(f"| `{n}` | {k[0]} | {k[1]/1e6:.0f} | {k[2]/1e6:.0f} | {k[3]:.0f} |\n")


# ---- full captures ------------------------------------------------------------------------------------------------
def capture(rep, title, out, want, ncols=2):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(title + "\n\n| metric | unit | " + " | ".join(f"launch {chr(65+i)}" for i in range(min(ncols, len(rows) - 2))) + " |\n|---|---|" + "---:|" * min(ncols, len(rows) - 2) + "\n")
        for i, name in enumerate(h):
            if name in want:
                f.write(f"| {name} | {units[i]} | " + " | ".join(r[i][:70] for r in rows[2:2 + ncols]) + " |\n")


WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
capture(os.path.join(G, f"gcfn_{tag}_final.ncu-rep"),
        f"# {tag} - `ncu --set full --clock-control none --import-source on -k regex:k_gcfn -s 57 -c 2` (tools/one_forward.py, B=32, Base)\n\n"
        "Two consecutive GCFN launches of the first encoder stage (32 x 8000 frames = 2752 tiles of 94 frames on 74 clusters of 2 CTAs).",
        os.path.join(OUT, f"{tag}_gcfn_ncu.md"), WANT)
capture(os.path.join(G, f"attn_{tag}.ncu-rep"),
        f"# {tag} - `ncu --set full --clock-control none --import-source on -k regex:k_attn_relpos -s 23 -c 1` (tools/one_forward.py, B=32, Base)\n\n"
        "One pooled-attention launch of the decoder half (64 rows x 8 heads x 500 keys).",
        os.path.join(OUT, f"{tag}_attn_ncu.md"), WANT, ncols=1)
print(open(os.path.join(OUT, f"{tag}_launches.md")).read()[:2200])
print(open(os.path.join(OUT, f"{tag}_forward_traffic.md")).read()[:1800])
print(open(os.path.join(OUT, f"{tag}_gcfn_ncu.md")).read())
