"""whvi_b200: B200-native (sm_100a) implementation of the WHVI hot path behind the
reference's Python surface.  See DESIGN.md for the path, INTEGRATION.md for the C ABI."""
from .fwht import FWHT, FWHTFunction, fwht_  # noqa: F401

__all__ = ["FWHT", "FWHTFunction", "fwht_"]
