from .CtrlAviary import CtrlAviary  # noqa: F401
from .HoverAviary import HoverAviary  # noqa: F401
from .MultiHoverAviary import MultiHoverAviary  # noqa: F401
from .VelocityAviary import VelocityAviary  # noqa: F401
