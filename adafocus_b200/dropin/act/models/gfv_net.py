"""Re-export of adafocus_b200.models.gfv_net under the reference's module path (models/gfv_net.py)."""
from adafocus_b200.models.gfv_net import *  # noqa: F401,F403
from adafocus_b200.models import gfv_net as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
