"""Drop-in for utils/volume.py of maxingan2412/Signal (B200 implementation; only
volume_computation3 is ever called by the reference, useB.py:106,110)."""
from signal_b200.modules import volume_computation3, volume_computation4, volume_computation5  # noqa: F401
