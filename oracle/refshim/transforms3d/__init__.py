"""Empty stand-in: only needed because the reference's envs/__init__.py imports BetaAviary."""
from . import quaternions, utils  # noqa: F401
