"""Drop-in for the evaluation entry points of utils/metrics.py of maxingan2412/Signal (B200 implementation):
euclidean_distance (:494), eval_func (:111), R1_mAP_eval (:222).  The plotting / t-SNE helpers of the reference file
are visualisation code outside the hot path and are not provided."""
from signal_b200.evaluation import R1_mAP_eval, euclidean_distance, eval_func  # noqa: F401
