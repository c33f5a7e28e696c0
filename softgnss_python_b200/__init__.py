"""B200-native GPS L1 C/A acquisition and tracking (drop-in for SoftGNSS-python's hot paths)."""
from .settings import Settings  # noqa: F401
