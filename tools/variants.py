"""Build and time compile-time variants of the library in one go (development aid for A/B measurements).

  python tools/variants.py build            # here (nvcc cross-compiles): sepreformer_b200/variants/libsepref_<name>.so
  python tools/variants.py time [B]         # on the GPU box: forward time of every built variant, same process layout

Variants are the -D switches of csrc/kernels_tc.cuh (poll-loop strategies).  The .so files are git-ignored but travel
with gpurun snapshots, so `gpurun -- python tools/variants.py time` measures them all in one call."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "sepreformer_b200", "variants")
VARIANTS = {
    "default": [],
    "suspend500": ["-DSEPREF_MBAR_SUSPEND_NS=500"],
    "suspend2000": ["-DSEPREF_MBAR_SUSPEND_NS=2000"],
    "fanout": ["-DSEPREF_FANOUT_WAITS"],
    "fanout_suspend2000": ["-DSEPREF_FANOUT_WAITS", "-DSEPREF_MBAR_SUSPEND_NS=2000"],
    "inline_waits": ["-DSEPREF_INLINE_WAITS"],       # profiling only: attributes polling to its wait site in ncu
}


def build():
    from __graft_entry__ import NVCC_FLAGS, CSRC
    os.makedirs(VDIR, exist_ok=True)
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, f"libsepref_{name}.so")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), *NVCC_FLAGS, *flags, "-o", out, os.path.join(CSRC, "sepref_api.cu")]
        print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def time(batch):
    for name in VARIANTS:
        so = os.path.join(VDIR, f"libsepref_{name}.so")
        if not os.path.exists(so):
            continue
        env = dict(os.environ, SEPREF_LIB=so)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_time.py"), "SepReformer_Base_WSJ0", str(batch), "2"],
                           env=env, capture_output=True, text=True, timeout=300)
        print(f"{name:22s}", (r.stdout.strip().splitlines() or [r.stderr.strip()[-200:]])[-1], flush=True)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    if cmd == "build":
        build()
    else:
        time(int(sys.argv[2]) if len(sys.argv) > 2 else 32)
