from .msml import MSML  # noqa: F401
from .fm import FMCnn, FMNone  # noqa: F401
from .osb import unet  # noqa: F401
from .frb import iresnet18, iresnet34, iresnet50  # noqa: F401
from ..headers.margin_losses import Softmax, AMCosFace, AMArcFace  # noqa: F401
