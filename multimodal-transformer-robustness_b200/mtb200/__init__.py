"""mtb200 -- B200-native (sm_100a) kernels and host runtime for the dynamic MulT
transformer stacks.  Importing this package loads libmultb200.so and fails loudly if it
has not been built (there is no CPU fallback)."""
from . import _lib  # noqa: F401  (loads the shared library)
from .ops import manual_seed, set_gemm_mode, get_gemm_mode, set_attn_mode  # noqa: F401

__all__ = ["manual_seed", "set_gemm_mode", "get_gemm_mode", "set_attn_mode"]
