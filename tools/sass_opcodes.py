"""Per-kernel SASS opcode evidence for the shipped library (profiles/r2_sass_opcodes.md):
counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and of the legacy tensor path (HMMA = mma.sync).

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sepreformer_b200", "libsepref_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", sass), filt))
cols = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "MUFU", "total"]
rows = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        short = re.sub(r"\(.*", "", names[m.group(1)]).replace("sepref::", "")
        short = re.sub(r"TokCfg<([^>]*)>", lambda g: "TokCfg<" + g.group(1).replace(" ", "") + ">", short)
        cur = rows.setdefault(short, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for c in cols:
            if op.startswith(c):
                cur[c] += 1
arch = re.findall(r"arch = (sm_\w+)", sass)
print(f"# r2 - SASS opcode summary of `sepreformer_b200/libsepref_b200.so` ({len(rows)} kernels, cubin arch {sorted(set(arch))})\n")
print("`cuobjdump -sass` of the library that ships to the GPU box, counted per kernel by `tools/sass_opcodes.py`.")
print("UTC*MMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, UTMALDG = TMA tensor load (`cp.async.bulk.tensor`), UBLKCP = `cp.async.bulk`,")
print("UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops, HMMA = legacy `mma.sync` (the attention kernel only), FFMA2 = packed `fma.rn.f32x2`.\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
tot = collections.Counter()
for k, c in rows.items():
    if not any(c[x] for x in cols[:-3]) and "k_" not in k:
        continue
    print(f"| `{k[:110]}` | " + " | ".join(str(c[x]) for x in cols) + " |")
    tot.update(c)
print("| **all kernels** | " + " | ".join(str(tot[x]) for x in cols) + " |")
