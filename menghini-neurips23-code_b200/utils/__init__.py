from .clip_pseudolabels import (compute_pseudo_labels, encode_pool, path_ranks, pseudolabel_top_k,
                                scan_features)

__all__ = ["compute_pseudo_labels", "encode_pool", "path_ranks", "pseudolabel_top_k", "scan_features"]
