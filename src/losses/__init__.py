"""Drop-in for the reference's src/losses (losses.py:10-47)."""
from .losses import InfoNCELoss

__all__ = ["InfoNCELoss"]
