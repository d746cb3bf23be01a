"""Drop-in for modeling/AddModule/useB.py of maxingan2412/Signal (B200 implementation)."""
from signal_b200.modules import AlignmentM  # noqa: F401
from signal_b200.modules import DA_sample as DAS  # noqa: F401  (useB.py:15 imports it under this name)
