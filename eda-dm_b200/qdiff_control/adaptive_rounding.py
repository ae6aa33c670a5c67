"""The reference keeps a copy of AdaRoundQuantizer here (qdiff_control/adaptive_rounding.py); one implementation serves both."""
from qdiff.adaptive_rounding import AdaRoundQuantizer  # noqa: F401
