"""Re-export of adafocus_b200.models_sth.ppo under the reference's module path (models/ppo.py)."""
from adafocus_b200.models_sth import ppo as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
