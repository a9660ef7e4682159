"""B200-native bundle adjustment / global positioning for InstantSfM (hot path only)."""
__version__ = "0.1.0"
