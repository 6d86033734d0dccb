"""The reference's training scripts run UNCHANGED on the B200 with `networks` = cenet_b200.

`baseline/_ref/src/main_{acdc,synapse,skin}.py` are byte copies of /root/reference/src (tools/vendor_reference.py, run by
__graft_entry__.build(); git-ignored, shipped to the GPU box).  Each test starts the script in a subprocess exactly as a user
would (`python -P main_x.py --flags`, PYTHONPATH = cenet_b200 : tests/stubs : src), on a small synthetic dataset in the
README layout, in the scripts' real mode: fp16 `autocast` + `GradScaler`, SGD momentum 0.9, poly schedule,
`--loss_type boundary` (ACDC / Synapse; `dice,ce` + AdamW for skin), followed by the per-volume B=1 evaluation, checkpoint
save, and a second `--eval` process that re-loads the checkpoint.  Asserted: exit code 0, finite loss that decreases,
checkpoint round trip (strict 801-key load) and identical Dice from the in-training evaluation and the `--eval` process.
"""
import glob
import math
import os
import re

import pytest
import torch

import mains_harness as H

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.have_reference(), reason="baseline/_ref not vendored")]


def _check_losses(save_dir, min_iters):
    logs = glob.glob(os.path.join(save_dir, "*", "log"))
    assert logs, os.listdir(save_dir)
    losses = H.read_scalars(logs[0])
    assert len(losses) >= min_iters, losses
    assert all(math.isfinite(v) for v in losses), losses
    k = max(1, len(losses) // 3)
    assert sum(losses[-k:]) / k < sum(losses[:k]) / k, losses          # the loss goes down on the (learnable) synthetic set
    return losses


def test_main_acdc_unchanged(tmp_path):
    d = str(tmp_path)
    ds = H.make_acdc(d, size=224, n_train=8, n_valid=2, n_vol=1, depth=3)
    base = ["--root_dir", ds["root_dir"], "--list_dir", ds["list_dir"], "--volume_path", ds["volume_path"],
            "--save_path", os.path.join(d, "out"), "--tag", "t", "--batch_size", "4", "--max_epochs", "4", "--base_lr", "0.01",
            "--img_size", "224", "--no_ptenc", "--amp", "--optimizer", "SGD", "--loss_type", "boundary",
            "--scale_factors", "1.0,0.5", "--num_heads", "4,4,4", "--out_up_block", "upcn"]        # scripts/acdc.sh:53-77
    r = H.run_main("main_acdc.py", base)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "AMP enabled" in r.stdout and "Using SGD optimizer" in r.stdout
    _check_losses(os.path.join(d, "out"), 8)
    ck = glob.glob(os.path.join(d, "out", "*", "best.pth"))
    assert ck, r.stdout[-2000:]
    sd = torch.load(ck[0], weights_only=True)
    assert len(sd) == 801 and all(torch.isfinite(v.float()).all() for v in sd.values())
    te = [float(m) for m in re.findall(r"te_DCS:([0-9.]+)", r.stdout)]
    r2 = H.run_main("main_acdc.py", base + ["--eval"])
    assert r2.returncode == 0, r2.stdout[-4000:]
    m = re.search(r"Average Dice: ([0-9.]+)", r2.stdout)
    assert m, r2.stdout[-2000:]
    # best.pth was written at the epoch with the best test Dice; the fresh process must reproduce that number
    assert abs(float(m.group(1)) * 100 - max(te)) < 0.02, (m.group(1), te)


def test_main_synapse_unchanged(tmp_path):
    d = str(tmp_path)
    ds = H.make_synapse(d, size=224, n_train=8, n_vol=1, depth=3)
    base = ["--root_dir", ds["root_dir"], "--list_dir", ds["list_dir"], "--volume_path", ds["volume_path"],
            "--save_path", os.path.join(d, "out"), "--tag", "t", "--batch_size", "4", "--max_epochs", "4", "--base_lr", "0.015",
            "--img_size", "224", "--no_ptenc", "--amp", "--optimizer", "SGD", "--loss_type", "boundary",
            "--scale_factors", "0.8,0.4", "--num_heads", "16,8,8", "--out_up_block", "upcn", "--num_workers", "0",
            "--eval_interval", "2"]                                                                 # scripts/synapse.sh:42-81
    r = H.run_main("main_synapse.py", base)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "Using CENet model" in r.stdout and "AMP enabled" in r.stdout
    _check_losses(os.path.join(d, "out"), 8)
    ck = sorted(glob.glob(os.path.join(d, "out", "*", "cenet_seed_1234_epoch_3.pth")))
    assert ck, r.stdout[-2000:]
    te = [float(m) for m in re.findall(r"te_DCS:([0-9.]+)", r.stdout)]
    r2 = H.run_main("main_synapse.py", base + ["--eval", "--checkpoint", ck[0]])
    assert r2.returncode == 0, r2.stdout[-4000:]
    m = re.search(r"Average Dice: ([0-9.]+)", r2.stdout)
    assert m and te, r2.stdout[-2000:]
    assert abs(float(m.group(1)) * 100 - te[-1]) < 0.02, (m.group(1), te)


def test_main_synapse_test_org_unchanged(tmp_path):
    """scripts/synapse.sh TEST_ORG: `main_synapse.py --model_version cenet_org --eval --checkpoint <published layout>` with
    `networks.CENetOrg` = cenet_b200 (822-key checkpoint, strict load, per-volume B=1 evaluation)."""
    from cenet_b200.networks import CENetOrg
    d = str(tmp_path)
    ds = H.make_synapse(d, size=224, n_train=2, n_vol=1, depth=3)
    torch.manual_seed(7)
    ck = os.path.join(d, "cenet_org.pth")
    torch.save(CENetOrg(num_classes=9, input_channels=1, scale_factors=[0.8, 0.4], encoder="pvt_v2_b2", pretrain=False,
                        num_heads=[16, 8, 8]).state_dict(), ck)
    args = ["--root_dir", ds["root_dir"], "--list_dir", ds["list_dir"], "--volume_path", ds["volume_path"],
            "--save_path", os.path.join(d, "out"), "--tag", "org", "--batch_size", "4", "--img_size", "224", "--num_workers", "0",
            "--model_version", "cenet_org", "--eval", "--checkpoint", ck]
    r = H.run_main("main_synapse.py", args)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "Using CENetOrg model" in r.stdout and re.search(r"Average Dice: ([0-9.]+)", r.stdout), r.stdout[-2000:]


def test_main_skin_unchanged(tmp_path):
    d = str(tmp_path)
    ds = H.make_ph2(os.path.join(d, "PH2"), size=224)
    base = ["--data_dir", ds["data_dir"], "--save_path", os.path.join(d, "out"), "--tag", "t", "--batch_size", "20",
            "--max_epochs", "2", "--base_lr", "0.0005", "--img_size", "224", "--no_ptenc", "--amp", "--optimizer", "AdamW",
            "--loss_type", "dice,ce", "--loss_weights", "0.7,0.3", "--scale_factors", "1.0,0.75,0.5", "--num_heads", "2,2,2"]
    r = H.run_main("main_skin.py", base)                                                           # scripts/skin.sh:45-100
    assert r.returncode == 0, r.stdout[-4000:]
    _check_losses(os.path.join(d, "out"), 8)
    assert glob.glob(os.path.join(d, "out", "*", "best.pth")), r.stdout[-2000:]
