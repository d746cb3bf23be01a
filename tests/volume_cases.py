"""Case table and seeded inputs of tests/golden/make_volume_golden.py (imported by path: generator and tests cannot drift)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_volume_golden", os.path.join(HERE, "golden", "make_volume_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)   # (touches /root/reference only inside main())
CASES = gen.CASES


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", f"volume_{name}.npz")))
