"""tsg -- B200-native hot path of two-stage-gnn graph-classification training.

Importing this package loads libtsg.so (hand-written sm_100a kernels behind a C ABI); there is
no CPU or eager fallback.  `tsg.synth` (input generation) is importable without the library.
"""
__all__ = ["ops", "nn", "synth"]
