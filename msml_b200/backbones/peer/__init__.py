from .arcface import arcface18, arcface34, arcface50, arcface100, cosface50_casia  # noqa: F401
