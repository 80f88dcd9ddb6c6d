"""B200-native batched quadrotor simulator: the ``Physics.DYN`` hot path of
komxun/gym-pybullet-drones-routing behind the reference's own Python API.

Import as ``gpd_b200`` (see ``gpd_b200.py`` at the repository root).  The compute path is the
C-ABI CUDA library ``lib/libgpd_b200.so`` (``include/gpd.h``); there is no CPU fallback.
"""
from .utils.enums import ActionType, DroneModel, ImageType, ObservationType, Physics  # noqa: F401

__version__ = "0.1.0"
