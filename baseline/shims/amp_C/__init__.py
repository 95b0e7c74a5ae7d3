"""Names the reference binds at import time (``optimization.py:29-32``) and never uses."""


def _unused(*a, **k):
    raise RuntimeError("amp_C stand-in: the reference never calls this")


multi_tensor_l2norm = multi_tensor_lamb_stage1_cuda = multi_tensor_lamb_stage2_cuda = multi_tensor_scale = _unused
