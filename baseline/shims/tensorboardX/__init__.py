"""Stub: the reference imports tensorboardX.SummaryWriter (rank-0 scalars only); not installed here."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None
