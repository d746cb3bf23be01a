"""signal_b200 -- B200-native fusion head (SIM + GAM + LAM) of Signal.

The CUDA extension is loaded lazily by ``signal_b200.lib``; helpers such as
``signal_b200.synthetic`` import without it.
"""
__version__ = "0.1.0"
