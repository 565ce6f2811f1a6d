from .dataloaderx import BackgroundGenerator, DataLoaderX  # noqa: F401
