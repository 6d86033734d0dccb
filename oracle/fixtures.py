"""TEST INFRASTRUCTURE ONLY -- deterministic weights / inputs shared by the golden generator, tests and bench.

The reference's init leaves most decoder branches numerically invisible (layer_scale = 1e-6, BN running stats
0/1, zero biases), so `perturb_state` rewrites those entries with seeded values (SURVEY.md section 8c recipe).
torch's CPU generator is bit-reproducible for a fixed torch build; the GPU boxes run the same image.
"""
from __future__ import annotations

import torch

CONFIGS = {
    # BASELINE.json configs -> constructor kwargs (scripts/acdc.sh:53-77, synapse.sh:42-81, skin.sh:45-100)
    "acdc": dict(input_channels=1, num_classes=4, scale_factors=[1.0, 0.5], diffatt_num_heads=[4, 4, 4],
                 out_up_block="upcn"),
    "synapse": dict(input_channels=1, num_classes=9, scale_factors=[0.8, 0.4], diffatt_num_heads=[16, 8, 8],
                    out_up_block="upcn"),
    "skin": dict(input_channels=3, num_classes=2, scale_factors=[1.0, 0.75, 0.5], diffatt_num_heads=[2, 2, 2],
                 out_up_block="upcn"),
    # the other PVTv2 variants (encoder.py:14-33) -- same kernels, other depths / MLP ratios
    "acdc_b1": dict(input_channels=1, num_classes=4, scale_factors=[1.0, 0.5], diffatt_num_heads=[4, 4, 4],
                    out_up_block="upcn", encoder="pvt_v2_b1"),
    # ... and the merge / up-block variants (dseb.py:155, out.py:58-64, blocks.py:188-204)
    "acdc_add": dict(input_channels=1, num_classes=4, scale_factors=[1.0, 0.5], diffatt_num_heads=[4, 4, 4],
                     out_up_block="upcn", skip_mode="add", out_merge_mode="add"),
    "synapse_uprb": dict(input_channels=1, num_classes=9, scale_factors=[0.8, 0.4], diffatt_num_heads=[16, 8, 8],
                         out_up_block="uprb", dec_up_block="uprb"),
    "acdc_uptc": dict(input_channels=1, num_classes=4, scale_factors=[1.0, 0.5], diffatt_num_heads=[4, 4, 4],
                      out_up_block="uptc", dec_up_block="uptc"),
    "acdc_b5": dict(input_channels=1, num_classes=4, scale_factors=[1.0, 0.5], diffatt_num_heads=[4, 4, 4],
                    out_up_block="upcn", encoder="pvt_v2_b5"),
}


def perturb_state(sd: dict, seed: int = 1234) -> dict:
    """Return a copy of `sd` with BN stats, layer scales and zero-initialised biases made non-trivial."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k].detach().clone()
        if k.endswith("running_mean"):
            v = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.5 + torch.rand(v.shape, generator=g)
        elif "layer_scale" in k:
            v = 0.5 + 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith(".bias") and v.ndim == 1:
            v = v + 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith(".weight") and v.ndim == 1:           # norm gains
            v = v * (1.0 + 0.1 * torch.randn(v.shape, generator=g))
        elif k == "out.out.1.conv.conv.weight":               # trained-like logit margins (tolerance realism)
            v = v * 20.0
        elif k.endswith("num_batches_tracked"):
            v = torch.tensor(7, dtype=torch.long)
        out[k] = v
    return {k: out[k] for k in sd.keys()}


def synth_input(name: str, batch: int, size: int = 224, seed: int = 0) -> torch.Tensor:
    """Synthetic slices shaped like each dataset's pre-processed input."""
    g = torch.Generator().manual_seed(seed)
    cin = CONFIGS[name]["input_channels"]
    name = name.split("_")[0]                    # variants (acdc_b1, ...) use the input statistics of their base config
    if name == "synapse":
        return (torch.randn(batch, cin, size, size, generator=g) * 0.5).clamp_(-1, 1)
    if name == "skin":
        return torch.rand(batch, cin, size, size, generator=g)
    return torch.randn(batch, cin, size, size, generator=g)
