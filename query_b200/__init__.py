"""query_b200 — B200-native filter + GROUP BY path behind the N1QL operator API.

All computation lives in libn1gpu.so (CUDA, sm_100a); this package is the ctypes binding plus the
torch.distributed plumbing for the multi-GPU merge.  There is no CPU fallback."""
from ._lib import Ineligible, N1GpuError, declared_symbols, lib  # noqa: F401
from .api import MISSING, Mailbox, Missing, Operator, Query, Result, Table, init, jit_stats, launch_count, set_segment_dir  # noqa: F401
