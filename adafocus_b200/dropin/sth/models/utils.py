"""Re-export of adafocus_b200.models_sth.utils under the reference's module path (models/utils.py)."""
from adafocus_b200.models_sth import utils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
