"""STH/models/utils.py mirror: get_patch (:44-58) is the same crop as in the ACT tree."""
from ..models.utils import get_patch, random_crop  # noqa: F401
