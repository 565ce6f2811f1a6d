from .margin_losses import Softmax, AMCosFace, AMArcFace, MarginSoftmax, ArcFace, CosFace  # noqa: F401
from .partial_fc import PartialFC  # noqa: F401
from .pfc_sgd import PartialFCSGD  # noqa: F401
