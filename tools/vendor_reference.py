"""Copies the reference's script tree (`/root/reference/{src,scripts,requirements.txt}`) to `baseline/_ref/` UNMODIFIED.

`baseline/_ref/` is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored, so the copy
travels to the GPU box, where /root/reference does not exist.  Used by
  * tests/test_gpu_mains.py  -- runs main_acdc.py / main_synapse.py / main_skin.py unchanged against `networks` = cenet_b200
  * bench.py --impl reference-gpu / the `eager` object -- times the reference module in PyTorch eager on the B200.
The reference is a script tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref` cannot work;
a byte copy is the install.  Run from __graft_entry__.build() whenever /root/reference is present.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CENET_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "src")):
        if verbose:
            print(f"[vendor_reference] {SRC} not present; keeping existing {DST}" if os.path.isdir(DST) else
                  f"[vendor_reference] {SRC} not present and no vendored copy")
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for sub in ("src", "scripts"):
        for dp, dn, fn in os.walk(os.path.join(SRC, sub)):
            dn[:] = [d for d in dn if d != "__pycache__"]
            rel = os.path.relpath(dp, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in fn:
                s, d = os.path.join(dp, f), os.path.join(DST, rel, f)
                if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                    shutil.copyfile(s, d)
                    n += 1
    for f in ("requirements.txt",):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        print(f"[vendor_reference] {DST} up to date ({n} files copied)")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
