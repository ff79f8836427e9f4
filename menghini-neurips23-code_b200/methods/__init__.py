"""Caller-side pieces of the hot path's boundary:

    training_strategies   re-creation of the file the reference's scrape lacks
                          (`methods/<paradigm>/training_strategies.py`, imported at methods/*/__init__.py:1)
    pseudolabels          one `assign_pseudo_labels` for the nine copies in methods/*/*_fpl.py

Both are written against the seam names (`import clip`, `from models import …`, `from utils import …`,
`from accelerate import Accelerator`), so they are imported by `dropin.install()` AFTER it has registered those."""
