"""Copy the parts of the reference that run its model (pure Python) into ``baseline/_ref/`` (git-ignored; travels to
the GPU box with gpurun snapshots, where /root/reference does not exist).

    python tools/install_reference.py            # build container only (needs /root/reference)

What is copied - unmodified - and why:
  models/<name>/{model.py, configs.yaml, modules/*.py}   the reference Model / Separator (bench.py --impl reference,
                                                         the reference-on-B200 leg, the install() drop-in tests)
  utils/decorators.py                                    imported by every reference module (needs loguru: in the image)
  utils/implements/criterions.py                         PIT_SISNRi, to pin the device-side metric
  sample_wav/sample_WSJ.wav                              BASELINE.json configs[0] input (147 KB)
Nothing here is product code and nothing under sepreformer_b200/ imports it.
"""
import os
import shutil
import sys

SRC = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(SRC):
        sys.exit(f"{SRC} not found: run this in the build container")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for model in sorted(os.listdir(os.path.join(SRC, "models"))):
        mdir = os.path.join(SRC, "models", model)
        if not os.path.isdir(mdir):
            continue
        for rel in ("model.py", "configs.yaml", "modules/module.py", "modules/network.py"):
            src = os.path.join(mdir, rel)
            if os.path.exists(src):
                dst = os.path.join(DST, "models", model, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copy2(src, dst)
                n += 1
    for rel in ("utils/decorators.py", "utils/implements/criterions.py", "sample_wav/sample_WSJ.wav", "LICENSE"):
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(SRC, rel), dst)
        n += 1
    print(f"copied {n} files into {DST}")


if __name__ == "__main__":
    main()
