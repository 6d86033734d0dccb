"""TEST INFRASTRUCTURE: runs the reference's UNCHANGED training scripts (`baseline/_ref/src/main_{acdc,synapse,skin}.py`, a
byte copy of /root/reference/src made by tools/vendor_reference.py) against `networks` = cenet_b200.

* synthetic datasets in the layouts the reference's dataset classes read (README.md:141-166; dataset_acdc.py:80-114,
  dataset_synapse.py:100-125, datasets/skin/dataset_ph2.py:118-240)
* `run_main(...)`: `python -P main_x.py <args>` with cwd = baseline/_ref/src and
  PYTHONPATH = <repo>/cenet_b200 : tests/stubs : baseline/_ref/src  (-P keeps the script's own directory -- which holds the
  reference's `networks/` -- off sys.path[0]; the third-party packages this image lacks come from tests/stubs)
* `cpu_dry_run=True` starts the same script through tests/cpu_main_driver.py: `.cuda()` becomes the identity and the C-ABI
  ops are replaced by the torch emulations of tests/fake_ops.py, so the host side of the drop-in (import path, autograd
  boundary, optimizer / GradScaler / checkpoint plumbing) is covered by `-m "not gpu"` here; the kernels are not.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")
STUBS = os.path.join(ROOT, "tests", "stubs")
DROPIN = os.path.join(ROOT, "cenet_b200")


def have_reference():
    return os.path.isfile(os.path.join(REF_SRC, "main_acdc.py"))


def _blobs(rng, n, H, W, ncls):
    """n images [H,W] float32 in [0,1] with class-coloured discs + the label maps (learnable: intensity encodes class)"""
    yy, xx = np.mgrid[0:H, 0:W]
    imgs, labs = [], []
    for _ in range(n):
        lab = np.zeros((H, W), np.uint8)
        for c in range(1, ncls):
            cy, cx = rng.integers(H // 6, 5 * H // 6), rng.integers(W // 6, 5 * W // 6)
            r = rng.integers(max(3, H // 12), max(4, H // 5))
            lab[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = c
        img = lab.astype(np.float32) / max(1, ncls - 1) * 0.8 + 0.1 + rng.normal(0, 0.03, (H, W)).astype(np.float32)
        imgs.append(np.clip(img, 0, 1).astype(np.float32))
        labs.append(lab)
    return np.stack(imgs), np.stack(labs)


def make_acdc(root, size=224, n_train=8, n_valid=2, n_vol=1, depth=3, vol_hw=(80, 72), seed=0):
    """ACDC: <root>/train|valid/<slice>.npz {img,label [H,W]}, <root>/test/<vol>.npz {img,label [D,H,W]}, lists_ACDC/*.txt"""
    rng = np.random.default_rng(seed)
    lists = os.path.join(root, "lists_ACDC")
    for d in ("train", "valid", "test", "lists_ACDC"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    for split, n in (("train", n_train), ("valid", n_valid)):
        x, y = _blobs(rng, n, size, size, 4)
        names = []
        for i in range(n):
            nm = f"case_{i:03d}_slice_0.npz"
            np.savez(os.path.join(root, split, nm), img=x[i], label=y[i])
            names.append(nm)
        open(os.path.join(lists, split + ".txt"), "w").write("\n".join(names) + "\n")
    names = []
    for v in range(n_vol):
        x, y = _blobs(rng, depth, vol_hw[0], vol_hw[1], 4)
        nm = f"vol_{v:03d}.npz"
        np.savez(os.path.join(root, "test", nm), img=x, label=y)
        names.append(nm)
    open(os.path.join(lists, "test.txt"), "w").write("\n".join(names) + "\n")
    return dict(root_dir=root, list_dir=lists, volume_path=os.path.join(root, "test"))


def make_synapse(root, size=224, n_train=8, n_vol=1, depth=3, vol_hw=(96, 80), seed=1):
    """Synapse: <root>/train_npz/<slice>.npz {image,label}, <root>/test_vol_h5/<case>.npy.h5 {image,label [D,H,W]} (stored as
    .npz archives under that name, read by the tests/stubs h5py stand-in), lists_Synapse/{train,test_vol}.txt"""
    rng = np.random.default_rng(seed)
    tr, te, lists = os.path.join(root, "train_npz"), os.path.join(root, "test_vol_h5"), os.path.join(root, "lists_Synapse")
    for d in (tr, te, lists):
        os.makedirs(d, exist_ok=True)
    x, y = _blobs(rng, n_train, size, size, 9)
    names = []
    for i in range(n_train):
        nm = f"case0005_slice{i:03d}"
        np.savez(os.path.join(tr, nm + ".npz"), image=x[i], label=y[i].astype(np.float32))
        names.append(nm)
    open(os.path.join(lists, "train.txt"), "w").write("\n".join(names) + "\n")
    names = []
    for v in range(n_vol):
        x, y = _blobs(rng, depth, vol_hw[0], vol_hw[1], 9)
        nm = f"case{v:04d}"
        with open(os.path.join(te, nm + ".npy.h5"), "wb") as fh:
            np.savez(fh, image=x, label=y.astype(np.float32))
        names.append(nm)
    open(os.path.join(lists, "test_vol.txt"), "w").write("\n".join(names) + "\n")
    return dict(root_dir=tr, list_dir=lists, volume_path=te)


def make_ph2(root, size=96, seed=2):
    """PH2: the pre-saved cache <root>/np/{X,Y}_tr_<S>x<S>.npy the reference loads (dataset_ph2.py:132-135,222-240):
    200 RGB images [200,3,S,S] in [0,1] and binary masks [200,1,S,S]; split 80 / 20 / 100 by the reference itself."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "np"), exist_ok=True)
    g, lab = _blobs(rng, 200, size, size, 2)
    X = np.stack([g, g * 0.7 + 0.1, 1.0 - g], 1).astype(np.float32)
    Y = lab[:, None].astype(np.float32)
    np.save(os.path.join(root, "np", f"X_tr_{size}x{size}.npy"), X)
    np.save(os.path.join(root, "np", f"Y_tr_{size}x{size}.npy"), Y)
    return dict(data_dir=root)


def run_main(script, args, cpu_dry_run=False, timeout=1500, env_extra=None):
    """Start baseline/_ref/src/<script> unchanged.  Returns CompletedProcess (stdout+stderr merged in .stdout)."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, STUBS, REF_SRC] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env.setdefault("MPLBACKEND", "Agg")
    if env_extra:
        env.update(env_extra)
    if cpu_dry_run:
        cmd = [sys.executable, "-P", os.path.join(ROOT, "tests", "cpu_main_driver.py"), os.path.join(REF_SRC, script)] + args
    else:
        cmd = [sys.executable, "-P", os.path.join(REF_SRC, script)] + args
    return subprocess.run(cmd, cwd=REF_SRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout)


def read_scalars(logdir, tag="info/criterion"):
    """losses written by the tests/stubs tensorboardX stand-in"""
    out = []
    p = os.path.join(logdir, "scalars.txt")
    if os.path.exists(p):
        for ln in open(p):
            t, step, v = ln.rstrip("\n").split("\t")
            if t == tag:
                out.append(float(v))
    return out
