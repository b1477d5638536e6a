"""Re-export of adafocus_b200.models.ppo under the reference's module path (models/ppo.py)."""
from adafocus_b200.models.ppo import *  # noqa: F401,F403
from adafocus_b200.models import ppo as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
