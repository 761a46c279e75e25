"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (run in the build container)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    return h, [r for r in rows[hi + 1:] if len(r) == len(h)]


def short(name):
    n = name.replace("void ", "").replace("sepref::", "")
    return n.split("(CUtensorMap")[0].split("(const")[0].split("(float")[0].split("(sepref")[0][:110]


# ---- launch list --------------------------------------------------------------------------------------------------
h, body = read_ncu_csv(os.path.join(G, f"launches_{tag}.csv"))
ci = {n: i for i, n in enumerate(h)}
acc = collections.OrderedDict()
tot = 0.0
for r in body:
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1000 if unit.startswith("n") else v if unit.startswith("u") else v * 1000
    a = acc.setdefault(short(r[ci["Kernel Name"]]), [0, 0.0])
    a[0] += 1; a[1] += us; tot += us
with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag} - ncu launch list of `bench.py --steps 2 --warmup 1` (2 timed forwards, B=32, Base)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -s 263 -c 526` - per-launch times are cold-cache and\n"
            "serialised, so compare SHARES with the CUDA-event numbers in the bench line, not absolutes.\n\n"
            f"{sum(a[0] for a in acc.values())} launches, {tot/1000:.2f} ms summed ({tot/2000:.2f} ms per forward).\n\n"
            "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us:.0f} | {100*us/tot:.1f} % |\n")
subprocess.run(["cp", os.path.join(G, f"launches_{tag}.csv"), os.path.join(OUT, f"{tag}_launches_raw.csv")], check=True)

# ---- GCFN dram traffic per launch ----------------------------------------------------------------------------------
h, body = read_ncu_csv(os.path.join(G, f"gcfn_dram_{tag}.csv"))
ci = {n: i for i, n in enumerate(h)}
per = collections.defaultdict(dict)
for r in body:
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]].lower()
    if "byte" in unit:
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        v *= mult
    elif unit.startswith("n"): v /= 1000
    elif unit.startswith("m") and "second" in unit: v *= 1000
    per[r[ci["ID"]]][r[ci["Metric Name"]]] = v
n = len(per)
rd = sum(p["dram__bytes_read.sum"] for p in per.values()); wr = sum(p["dram__bytes_write.sum"] for p in per.values())
us = sum(p["gpu__time_duration.sum"] for p in per.values())
B, Tp, F = 32, 8000, 128
alg = 2 * 4 * F * 41.5 * B * Tp            # x in + y out, fp32, 41.5 token-calls per padded frame
json.dump({"launches": n, "dram_bytes_per_launch": (rd + wr) / n, "dram_read_bytes_per_forward": rd, "dram_write_bytes_per_forward": wr,
           "algorithmic_bytes_per_forward": alg, "algorithmic_bytes_per_launch": alg / n, "ncu_time_us_per_forward": us},
          open(os.path.join(OUT, f"{tag}_gcfn_traffic.json"), "w"), indent=1)

# ---- GCFN full capture: key metrics ---------------------------------------------------------------------------------
rep = os.path.join(G, f"gcfn_{tag}_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_op_utcmma.sum" ,
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
with open(os.path.join(OUT, f"{tag}_gcfn_ncu.md"), "w") as f:
    f.write(f"# {tag} - `ncu --set full --clock-control none --import-source on -k regex:k_gcfn -s 57 -c 2` (one_forward, B=32, Base)\n\n"
            "Two consecutive GCFN launches of the first encoder stage (32 x 8000 frames, 2752 tiles of 96 frames, 74 clusters of 2).\n\n"
            "| metric | unit | launch A | launch B |\n|---|---|---:|---:|\n")
    for i, name in enumerate(h):
        if name in want:
            vals = [r[i][:70] for r in rows[2:4]]
            f.write(f"| {name} | {units[i]} | " + " | ".join(vals) + " |\n")
    f.write(f"\nDRAM traffic over the 56 GCFN launches of one forward (`{tag}_gcfn_traffic.json`): read {rd/1e9:.2f} GB + write {wr/1e9:.2f} GB = "
            f"{(rd+wr)/1e9:.2f} GB against {alg/1e9:.2f} GB algorithmic (x in + y out): ratio {(rd+wr)/alg:.2f}.\n")
print(open(os.path.join(OUT, f"{tag}_launches.md")).read()[:2500])
print(open(os.path.join(OUT, f"{tag}_gcfn_ncu.md")).read())
