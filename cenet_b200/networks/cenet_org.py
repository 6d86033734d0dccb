"""`networks.CENetOrg` (reference: src/networks/cenet_org/net.py:16-129).

main_synapse.py:15 imports the symbol, so it must exist; the "original paper" variant itself is only reachable
through `synapse.sh TEST_ORG` and is a "next" row of the scope table, not part of
the accelerated hot path yet.  Constructing it fails loudly instead of silently falling back to another model.
"""
import torch.nn as nn


class CENetOrg(nn.Module):
    def __init__(self, num_classes=1, input_channels=1, scale_factors=[0.6, 0.3], num_heads=[2, 2, 2],
                 encoder="pvt_v2_b2", pretrain=False, skip_mode="cat", base_ptdir="."):
        super().__init__()
        raise NotImplementedError(
            "CENetOrg (cenet_org variant) has no sm_100a path yet; use networks.CENet "
            "(SURVEY.md section 8f rank 2)")
