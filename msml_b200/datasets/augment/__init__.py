from .rand_occ import RandomBlock, NoneOcc  # noqa: F401
