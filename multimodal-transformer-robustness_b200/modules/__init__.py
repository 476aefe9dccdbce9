"""Drop-in replacement of the reference's ``modules`` package: same module paths, class names,
constructor signatures, attribute names and state_dict keys, with the
math running in libmultb200's sm_100a kernels.  Put this package's parent directory ahead of
the reference on sys.path and the reference's src/dynamic_models2.py, src/train.py, main.py
and EA.py pick it up unchanged."""
from .dynamic_multihead_attention import *  # noqa: F401,F403
from .multihead_attention import *  # noqa: F401,F403
from .transformer import *  # noqa: F401,F403
