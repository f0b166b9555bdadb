"""
tapqir_b200 -- B200-native (sm_100a) implementation of Tapqir's cosmos SVI hot path.

Layout mirrors the reference modules that sit on that path (``tapqir.distributions``,
``tapqir.models``, ``tapqir.utils.dataset``); the arithmetic lives in ``csrc/`` behind the C ABI of
``include/tapqir_b200.h``.
"""

__version__ = "0.1.0"
