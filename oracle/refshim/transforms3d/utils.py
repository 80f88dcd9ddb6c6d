"""Stand-in names only (CTBRControl imports it; out of scope and never called)."""
def normalized_vector(*a, **k):
    raise NotImplementedError("transforms3d stand-in: out-of-scope reference path")
