"""Drop-in for modeling/AddModule/DAS.py of maxingan2412/Signal (B200 implementation)."""
from signal_b200.modules import DA_sample  # noqa: F401
