from .iresnet import IResNet, iresnet18, iresnet34, iresnet50  # noqa: F401
