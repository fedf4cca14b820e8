"""tf2_b200 — B200-native implementation of TF2's Runtime_Engine/cnn quantised-convolution path."""
from .netdesc import NetDesc, LayerDesc, TensorDesc  # noqa: F401
from .header_tables import parse_header, parse_header_file  # noqa: F401

__all__ = ["NetDesc", "LayerDesc", "TensorDesc", "parse_header", "parse_header_file"]
