"""TEST INFRASTRUCTURE ONLY -- `dgl.nn` stand-in namespace."""
from . import pytorch  # noqa: F401
