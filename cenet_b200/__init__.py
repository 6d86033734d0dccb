"""cenet_b200 -- B200-native (sm_100a) implementation of the CENet forward hot path.

Public surface:
  cenet_b200.networks.CENet / CENetOrg   drop-in for the reference's `networks` package
  cenet_b200.ops                         thin tensor -> C-ABI wrappers (one per kernel family)
  cenet_b200.engine.Engine               launch plan of one forward pass
  cenet_b200.criterion.DiceCELoss        fused Dice+CE (utils/core.py Criterion, 'dice,ce')
"""
__version__ = "0.1.0"
