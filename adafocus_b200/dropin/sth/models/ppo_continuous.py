"""Re-export of adafocus_b200.models_sth.ppo_continuous under the reference's module path (models/ppo_continuous.py)."""
from adafocus_b200.models_sth import ppo_continuous as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
