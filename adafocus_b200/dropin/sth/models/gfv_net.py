"""Re-export of adafocus_b200.models_sth.gfv_net under the reference's module path (models/gfv_net.py)."""
from adafocus_b200.models_sth import gfv_net as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
