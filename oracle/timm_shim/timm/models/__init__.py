from . import layers, registry, twins, vision_transformer  # noqa: F401
