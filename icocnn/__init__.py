"""Drop-in package name the reference imports (models.py:4-6, losses.py:4-5): binds to the B200
implementation in geniconet_b200.  `from icocnn.ico_conv import IcoConvS2S, IcoUpsampleS2S` works
unchanged once the repository root is on sys.path / PYTHONPATH."""
from . import ico_conv, utils  # noqa: F401
