"""grapes_b200 -- B200-native (sm_100a) implementation of the GRAPES per-batch
sampling-and-aggregation hot path (reference: dfdazac/grapes main.py:157-291, modules/gcn.py,
modules/utils.py), behind the reference's own Python surface."""
__version__ = "0.1.0"
