"""Drop-in for modeling/AddModule/useA.py of maxingan2412/Signal (B200 implementation)."""
from signal_b200.modules import LayerNorm, ModalInteractive, Select_Interactive_Module, TokenSelection  # noqa: F401
