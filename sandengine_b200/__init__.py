"""sandengine_b200 -- B200-native (sm_100a) falling-sand simulation core.

Drop-in for the ONE hot path of ARez2/sandengine: the Margolus 2x2 block update generated from
materials.yaml (+ modification override + flood-fill lighting).  The product is the native library
(libsandengine_b200.so, C ABI in include/sandengine_b200.h); this package is its Python host mirror of
the reference's `sandengine_lang` / `sandengine_core::simulation` interfaces.
"""
from . import _capi  # noqa: F401
from .lang import ParsingResult, SandEngineError, SandMaterial, SandRule, parse_path, parse_string, create_cuda_from_parser  # noqa: F401
from .simulation import (MAX_MODIFICATIONS, MOD_DTYPE, MODSHAPE_CIRCLE, MODSHAPE_SQUARE, Params, SimModification,  # noqa: F401
                         Simulation)

__version__ = "0.1.0"
