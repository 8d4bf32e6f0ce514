"""htnorm_b200: B200-native (sm_100a) sampler for hyperplane-truncated and structured-precision
multivariate normals - a drop-in for zoj613/htnorm's sampling path plus batched entry points.

The product is the C-ABI library ``libhtnorm_b200.so`` (C host layer + hand-written CUDA kernels);
this package is a thin ctypes mirror of the reference's Python interface.  Importing it without the
built library raises ImportError: there is no CPU fallback.
"""
from ._lib import LIB_PATH, PUBLIC_SYMBOLS, lib  # noqa: F401
from .api import (HTNGenerator, fp64_fma_peak_tflops, fp64_tensor_peak_tflops,  # noqa: F401
                  hyperplane_truncated_mvnorm, hyperplane_truncated_mvnorm_batch, int_issue_peak_gops, launch_count,
                  phase_timing,
                  raw_streams, standard_normals, structured_precision_mvnorm,
                  structured_precision_mvnorm_batch)
from .dist import PeerGather, gather_rows, sample_sharded, shard_range, shard_sizes  # noqa: F401

__version__ = "0.1.0"
