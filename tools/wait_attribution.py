"""Where do the warps of a kernel wait?  Post-processes `ncu -i <rep> --page source --csv --print-source cuda` (capture
taken with the `inline_waits` variant of tools/variants.py, so every mbarrier poll loop belongs to its call site):
prints the CUDA source lines with the most stall samples and their share.

  SEPREF_LIB=sepreformer_b200/variants/libsepref_inline_waits.so ncu --set full --import-source on -k regex:k_gcfn \\
      -s 57 -c 1 -o gpurun_out/gcfn_waits python tools/one_forward.py
  ncu -i gpurun_out/gcfn_waits.ncu-rep --page source --csv --print-source cuda > /tmp/src.csv
  python tools/wait_attribution.py /tmp/src.csv"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line", "#"))
h, body = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ci = {n: i for i, n in enumerate(h)}
samp = ci.get("Warp Stall Sampling (All Samples)", ci.get("# Samples"))
src = ci["Source"]
tot = sum(int(r[samp] or 0) for r in body)
print(f"{tot} stall samples over {len(body)} source lines")
for r in sorted(body, key=lambda r: -int(r[samp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    n = int(r[samp] or 0)
    print(f"{100.0 * n / max(tot, 1):5.1f} %  {n:6d}  {r[src].strip()[:150]}")
