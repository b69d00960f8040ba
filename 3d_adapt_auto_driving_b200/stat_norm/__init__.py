"""Statistical Normalization (stat_norm/ of the reference): rescale the points inside every Car / Van
box and the labels by the difference of the mean car size between two domains.
Re-exports the reference's public names (stat_norm/__init__.py:2)."""
from .norm import convert, launch_rescale, rescale_ptc, scale_labels, get_scale_map, single_scale, format_lidar_data  # noqa: F401
