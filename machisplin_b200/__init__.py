"""machisplin_b200 - B200-native TPS + ensemble raster interpolation behind the machisplin API.

The compute lives in ``libmachisplin_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/machisplin_b200.h``); this package is the host-side mirror of the reference's R interface
(``machisplin.mltps``, ``machisplin.tiles.create``, ``machisplin.tiles.merge``) plus the loader.
There is no CPU fallback: creating an Engine without the library or without a GPU raises.
"""
from .engine import Engine, Geom, Spline, Ensemble, as_geom, EVAL_DIRECT, EVAL_FAST  # noqa: F401
from .mltps import mltps_response, select_models, rss_objective_from_gram, knot_cells  # noqa: F401
from .tiles import tiles_create, tiles_merge, crop_window  # noqa: F401
from .geotiff import raster_info, read_raster, read_stack, write_raster, write_geotiff  # noqa: F401

__version__ = "0.1.0"
