from .clip_pseudolabels import (compute_pseudo_labels, encode_pool, path_ranks, pseudolabel_top_k,
                                scan_features)
from .evaluation import predict_features, predictions_frame, test_predictions
from .misc import Config, seed_worker
from .schedulers import WarmupCosineSchedule, make_scheduler

__all__ = ["compute_pseudo_labels", "encode_pool", "path_ranks", "pseudolabel_top_k", "scan_features",
           "predict_features", "predictions_frame", "test_predictions", "Config", "seed_worker",
           "WarmupCosineSchedule", "make_scheduler"]
