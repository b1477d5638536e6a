"""Re-export of adafocus_b200.models_sth.tsn under the reference's module path (models/tsn.py)."""
from adafocus_b200.models_sth import tsn as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
