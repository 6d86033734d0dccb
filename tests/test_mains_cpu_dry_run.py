"""Host side of the drop-in on CPU: the reference's UNCHANGED main_acdc.py (baseline/_ref/src, vendored copy) runs one epoch
+ per-volume evaluation + checkpoint save + a second `--eval` process with `networks` = cenet_b200, `.cuda()` mapped to the
identity and the C-ABI ops emulated in torch (tests/cpu_main_driver.py).  Covers: `from networks import CENet` via PYTHONPATH
(+ `python -P`), FlopCountAnalysis (= torch.jit.trace) and thop on a deepcopy, the 5 grad-enabled eval warm-ups
(utils/utils.py:171-185), train-mode autograd boundary under `autocast` + `GradScaler` + SGD, eval under grad mode (`val()`,
main_acdc.py:226), state_dict save / strict load.  The same scripts run on the real kernels in tests/test_gpu_mains.py."""
import glob
import math
import os
import re

import pytest

import mains_harness as H

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="baseline/_ref not vendored (tools/vendor_reference.py)")


def test_main_acdc_cpu_dry_run(tmp_path):
    d = str(tmp_path)
    S = 64
    ds = H.make_acdc(d, size=S, n_train=4, n_valid=1, n_vol=1, depth=2, vol_hw=(40, 36))
    base = ["--root_dir", ds["root_dir"], "--list_dir", ds["list_dir"], "--volume_path", ds["volume_path"],
            "--save_path", os.path.join(d, "out"), "--tag", "t", "--batch_size", "2", "--max_epochs", "1", "--base_lr", "0.01",
            "--img_size", str(S), "--no_ptenc", "--amp", "--scale_factors", "1.0,0.5", "--num_heads", "4,4,4",
            "--out_up_block", "upcn"]
    r = H.run_main("main_acdc.py", base, cpu_dry_run=True)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "Model parameters: 33384872" in r.stdout
    losses = H.read_scalars(glob.glob(os.path.join(d, "out", "*", "log"))[0])
    assert len(losses) == 2 and all(math.isfinite(v) for v in losses)
    assert glob.glob(os.path.join(d, "out", "*", "best.pth"))
    te = float(re.search(r"te_DCS:([0-9.]+)", r.stdout).group(1))
    r2 = H.run_main("main_acdc.py", base + ["--eval"], cpu_dry_run=True)
    assert r2.returncode == 0, r2.stdout[-4000:]
    assert abs(float(re.search(r"Average Dice: ([0-9.]+)", r2.stdout).group(1)) * 100 - te) < 0.02
