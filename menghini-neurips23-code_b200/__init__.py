"""grip-b200: the CLIP-encoder + soft-prompt + pseudolabel hot path of
BatsResearch/menghini-neurips23-code on B200 (sm_100a) — hand-written CUDA behind the reference's own
Python classes.  The directory name has hyphens; import it with
`importlib.import_module("menghini-neurips23-code_b200")`.

    .clip    stand-in for the third-party `clip` package (load / tokenize / model.Transformer)
    .models  CustomTextEncoder, CustomImageEncoder, TextEncoder, ImageEncoder,
             TextPrefixModel, ImagePrefixModel, UPTModel        (reference: models/)
    .utils   pseudolabel_top_k, compute_pseudo_labels            (reference: utils/clip_pseudolabels.py)
    .engine  Engine / Leaderboard: torch-level view of the C ABI (include/gripb200.h)
"""
from ._lib import Context, GripB200Error, LIB_PATH, declared_symbols, load  # noqa: F401

__version__ = "0.1"
