"""gcmapexplorer_b200 -- B200-native drop-in for the normalization hot path of gcMapExplorer.

Module names mirror ``gcMapExplorer.lib`` for the path: ``normalizer`` (outer seam), ``normalizeCore``,
``normalizeKnightRuiz``, ``ccmapHelpers``, ``cmstats`` (inner seam) and ``ccmap`` (the CCMAP carrier).
All numeric work runs in hand-written sm_100a CUDA kernels behind the C ABI of ``libgcxnorm.so``
(include/gcx_norm.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
