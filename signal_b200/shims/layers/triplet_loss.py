"""Drop-in for layers/triplet_loss.py of maxingan2412/Signal (B200 implementation)."""
from signal_b200.losses import TripletLoss, normalize  # noqa: F401
