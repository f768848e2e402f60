"""Run-time switches of the host layer.

strict_asserts: the reference validates its inputs with Python asserts that read device data back
(pseudo_generation.py:71 range assert = two host syncs; F.one_hot's value check inside DownscaleLabel /
_index2onehot).  With strict_asserts=True (default) the drop-in functions raise the same
AssertionError / RuntimeError on bad input, at the cost of one small device->host read per call.
Set to False in a tuned training loop to keep the stream free of host syncs.
"""
strict_asserts = True
