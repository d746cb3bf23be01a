"""Drop-in for layers/softmax_loss.py of maxingan2412/Signal (B200 implementation)."""
from signal_b200.losses import CrossEntropyLabelSmooth, LabelSmoothingCrossEntropy  # noqa: F401
