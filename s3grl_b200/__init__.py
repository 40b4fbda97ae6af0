"""s3grl_b200 — B200-native (sm_100a) precompute hot path of S3GRL.

Public surface (mirrors the reference for this path):
    extract_enclosing_subgraphs            reference utils.py:446
    OptimizedSignOperations                reference tuned_SIGN.py:47
    DeviceGraph, precompute                tensor-level API beneath them
    SEALDataset, get_pos_neg_edges         reference sgrl_link_pred.py:54, utils.py:637 (host orchestration)
    JointLoader, save_collated             GPU-resident loader / collated .pt (sgrl_link_pred.py:204, :1253)
The compute lives in lib/libs3grl_b200.so (include/s3grl_b200.h); build it with
`python -m s3grl_b200.build`.  Importing the package does not need a GPU; calling it does.
"""
__version__ = '0.1.0'

from .data import Data, PrecomputedList  # noqa: F401
from .engine import DeviceGraph, PrecomputeResult, algorithmic_bytes, pool_rows, precompute, precompute_full, walk_sets  # noqa: F401
from .tuned_sign import OptimizedSignOperations  # noqa: F401
from .utils import extract_enclosing_subgraphs  # noqa: F401
from .loader import JointLoader, joint_rows, load_collated, save_collated  # noqa: F401
from .dataset import SEALDataset, do_edge_split, do_edge_split_gpu, get_pos_neg_edges, sample_negative_edges_gpu  # noqa: F401
from .head import fold_batchnorm, segment_pool, sign_head, sign_head_ccn  # noqa: F401
