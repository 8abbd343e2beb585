"""Import-path compatibility: ``cs_vit.net.ti_poser`` is where the reference defines these classes."""
from .poser import PerspectiveEncoder, Poser, SpatialEncoder, TemporalEncoder, derivative  # noqa: F401
