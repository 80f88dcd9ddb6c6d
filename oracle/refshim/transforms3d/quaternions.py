"""Stand-in names only (BetaAviary/CTBRControl import them; both are out of scope and never called)."""
def _unavailable(*a, **k):
    raise NotImplementedError("transforms3d stand-in: out-of-scope reference path")
rotate_vector = qconjugate = mat2quat = qmult = _unavailable
