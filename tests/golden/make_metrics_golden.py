"""Golden vectors of the inference tail from the LIVE reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_metrics_golden.py

Imports /root/reference/utils/metrics.py (its plotting imports -- matplotlib, seaborn, PIL, scipy.integrate.simps -- are
stubbed in sys.modules: ``eval_func`` and ``euclidean_distance`` do not touch them) and stores, per case, the inputs'
seeds and the reference's distance matrix, CMC curve and mAP in tests/golden/metrics_<case>.npz.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True
import metrics_cases as mc  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.patches", "matplotlib.pyplot", "seaborn", "PIL", "PIL.Image"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    import scipy.integrate
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
    pkg = types.ModuleType("utils")
    pkg.__path__ = ["/root/reference/utils"]
    sys.modules["utils"] = pkg
    sys.path.insert(0, "/root/reference")
    from utils import metrics
    return metrics


if __name__ == "__main__":
    ref = import_reference()
    for name, c in mc.CASES.items():
        feats, pids, camids = mc.make_case(c)
        nq = c["nq"]
        f = torch.nn.functional.normalize(torch.from_numpy(feats), dim=1, p=2) if c["norm"] else torch.from_numpy(feats)
        dist = ref.euclidean_distance(f[:nq], f[nq:])
        cmc, mAP = ref.eval_func(dist, pids[:nq], pids[nq:], camids[:nq], camids[nq:], max_rank=c["max_rank"])
        np.savez_compressed(os.path.join(HERE, f"metrics_{name}.npz"), distmat=dist, cmc=cmc, mAP=np.float64(mAP))
        print(f"{name}: nq={nq} ng={feats.shape[0] - nq} mAP={mAP:.6f} R1={cmc[0]:.4f} R5={cmc[min(4, len(cmc) - 1)]:.4f}")
