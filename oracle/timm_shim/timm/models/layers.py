"""timm.models.layers (deprecated alias of timm.layers in 0.9)."""
from ..layers import DropPath, Mlp, drop_path, to_2tuple, trunc_normal_  # noqa: F401
