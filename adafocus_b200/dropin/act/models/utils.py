"""Re-export of adafocus_b200.models.utils under the reference's module path (models/utils.py)."""
from adafocus_b200.models.utils import *  # noqa: F401,F403
from adafocus_b200.models import utils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
