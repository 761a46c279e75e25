"""Scan the shipped library's SASS for global / generic loads whose first consumer follows within a few instructions with
no other load in between - the pattern that serialises a batch of loads (each one waits out the previous one's latency).

    python tools/sass_load_use.py [library.so] [min_hits]

Found with it in round 2: the raw-stream range tracking (an abs-max next to each load: k_tok<fuse> 0.35 -> 1.7 ms until it was
moved behind the loads) and the re-scheduled spk_proj<256> (16 residual loads each followed by its add: 2.7 -> 5.4 ms per
Large forward).  Remaining hits in the shipped library are the shared-memory residual reads of the gate epilogue (latency of
a shared-memory load, off the critical path of a producer-bound kernel) and the small SIMT kernels that rely on occupancy."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sepreformer_b200", "libsepref_b200.so")
min_hits = int(sys.argv[2]) if len(sys.argv) > 2 else 4
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for l in out.splitlines():
    if "Function :" in l:
        cur = l.split("Function :")[1].strip(); kern[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", l):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", l)
        if m: kern[cur].append(m.group(1).strip())
LOAD = re.compile(r"(@!?U?P\d+\s+)?(LDG\S*|LD\.E\S*)\s+R(\d+)")
rows = []
for name, L in kern.items():
    hits = tot = 0; ex = None
    for i, ins in enumerate(L):
        m = LOAD.match(ins)
        if not m: continue
        tot += 1
        w = 4 if ".128" in m.group(2) else 2 if ".64" in m.group(2) else 1
        dst = set(range(int(m.group(3)), int(m.group(3)) + w))
        for j in range(i + 1, min(i + 60, len(L))):
            if LOAD.match(L[j]): continue
            if dst & set(int(x) for x in re.findall(r"\bR(\d+)\b", L[j])):
                between = sum(1 for k in range(i + 1, j) if LOAD.match(L[k]))
                if j - i <= 8 and between <= 1:
                    hits += 1
                    ex = ex or (ins[:50], L[j][:44], j - i)
                break
    if hits >= min_hits: rows.append((hits, tot, name, ex))
print(f"{lib}: {len(kern)} kernels; kernels with >= {min_hits} loads consumed within 8 instructions (<= 1 load in between):")
for hits, tot, name, ex in sorted(rows, reverse=True):
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110] or name[:110]
    print(f"{hits:4d} of {tot:4d} loads  {short}\n        e.g. {ex[0]}  ->  {ex[1]}  (+{ex[2]})")
