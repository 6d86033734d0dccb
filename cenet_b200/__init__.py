"""cenet_b200 -- B200-native (sm_100a) implementation of the CENet forward hot path.

Public surface:
  cenet_b200.networks.CENet / CENetOrg   drop-in for the reference's `networks` package
  cenet_b200.ops                         thin tensor -> C-ABI wrappers (one per kernel family)
  cenet_b200.engine.Engine               launch plan of one forward pass
  cenet_b200.train.TrainEngine           launch plan of one training step (forward, fused loss, backward, AdamW)
  cenet_b200.losses.Criterion            drop-in for utils/core.py Criterion (fused Dice + CE + Boundary-DoU)
  cenet_b200.volume                      batched per-volume inference with on-device integer Dice counts
  cenet_b200.replicas                    batch sharding + bucketed gradient all-reduce (one process per GPU)
"""
__version__ = "0.1.0"
