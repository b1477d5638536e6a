"""Re-export of adafocus_b200.models.mobilenet under the reference's module path (models/mobilenet.py)."""
from adafocus_b200.models.mobilenet import *  # noqa: F401,F403
from adafocus_b200.models import mobilenet as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
