"""Re-export of adafocus_b200.models.resnet under the reference's module path (models/resnet.py)."""
from adafocus_b200.models.resnet import *  # noqa: F401,F403
from adafocus_b200.models import resnet as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
