"""jaxovercooked_b200 -- B200-native batched Overcooked environment step."""
from .layouts import overcooked_layouts, layout_grid_to_dict, pad_layouts  # noqa: F401

__version__ = "0.1.0"
