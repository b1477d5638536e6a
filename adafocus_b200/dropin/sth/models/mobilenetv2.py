"""Re-export of adafocus_b200.models_sth.mobilenetv2 under the reference's module path (models/mobilenetv2.py)."""
from adafocus_b200.models_sth import mobilenetv2 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
