"""Host-side mirror of the reference's `models` package (ACT tree): same module / function names so that
`from models.gfv_net import GFV` resolves here when adafocus_b200/dropin/act is put on sys.path."""
from .gfv_net import GFV, Focuser, Glancer, PatchSampler, RecurrentClassifier  # noqa: F401
from .utils import get_patch  # noqa: F401
