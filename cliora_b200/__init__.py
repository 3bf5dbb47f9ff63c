"""cliora_b200: B200-native (sm_100a) implementation of CLIORA's chart hot path.

Host-side mirror of the reference's ``cliora/net`` module API over a C-ABI CUDA
library (``include/cliora_b200.h``).  See DESIGN.md / INTEGRATION.md.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
__version__ = '0.1.0'
