"""blackstar_b200 -- B200-native (sm_100a CUDA) implementation of the hot path of
flannelhead/blackstar, behind the C ABI of include/blackstar_b200.h.

Host-side mirror of the reference's interface for this path:
  config   -- ConfigFile.hs / Main.prepareScene
  starmap  -- StarMap.hs (catalogue reader, spectral colours)
  render   -- Raytracer.render / ImageFilters.bloom / Raytracer.writeImg / Main.doRender
"""
from . import config, starmap  # noqa: F401

__all__ = ["config", "starmap", "render"]
__version__ = "0.1.0"
