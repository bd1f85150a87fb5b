"""Drop-in for the reference's `mesh.utils` import (losses.py:7, generate.py:13)."""
from . import utils  # noqa: F401
