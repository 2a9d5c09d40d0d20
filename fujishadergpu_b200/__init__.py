"""fujishadergpu_b200 -- B200-native (sm_100a) terrain-shading hot path behind the
geoign/FujiShaderGPU algorithm API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
