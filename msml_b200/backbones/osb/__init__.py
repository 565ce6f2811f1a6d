from .unet import unet, Unet  # noqa: F401
