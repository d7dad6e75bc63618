"""yak_b200 - B200-native k-mer counting / lookup behind lh3/yak's C API.

The product is the C-ABI shared library ``yak_b200/lib/libyakb200.so`` (sources in ``csrc/``,
interface in ``include/yak.h`` + ``include/yak_b200.h``); this package holds its ctypes binding
(`capi`) and the seeded synthetic-read generator (`synth`).
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
