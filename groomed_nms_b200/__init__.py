"""groomed_nms_b200 -- B200-native (sm_100a) implementation of the GrooMeD-NMS hot path of abhi1kumar/groomed_nms:
pairwise 2D / approximate-3D overlaps, score-sorted greedy grouping, grouped pruning-matrix rescore with its
analytic backward, the classical hard/soft NMS family and AP loss -- hand-written CUDA behind a C-ABI shared
library (include/groomed_nms_b200.h), exposed through the reference's own Python call surface
(groomed_nms_b200.lib.*).  There is no CPU fallback: every entry point raises if the CUDA library is missing."""
from .install import install  # noqa: F401

__version__ = "0.1.0"
