"""Host-side mirror of the reference's `models` package of the Something-Something tree (plus the two `ops` modules
the model code depends on: temporal_shift, basic_ops)."""
from .gfv_net import GFV, Focuser, Glancer  # noqa: F401
from .utils import get_patch  # noqa: F401
