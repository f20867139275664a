"""hcore_b200 -- B200-native tile-low-rank (TLR) update path of ecrc/hcore.

The product is the C-ABI shared library ``hcore_b200/lib/libhcore_b200.so``
(``include/hcore_b200.h``, hand-written sm_100a CUDA in ``hcore_b200/csrc``).  This
package is the thin host-side mirror used by tests and ``bench.py``:

* ``hcore_b200.api``       -- ``Context`` (batched, device tensors), the legacy single-tile
  calls ``hcore_dgemm`` / ``hcore_dgemmcd`` / ``hcore_dsyrk`` (host arrays, same argument
  meaning as the reference's C functions) and ``FlopCounter``
* ``hcore_b200.scheduler`` -- tile scheduler: shard a batch over ranks by tile index,
  NCCL (torch.distributed) gather of per-tile rank/status
* ``hcore_b200.tilegen``   -- synthetic TLR batches

PyTorch is plumbing only (device memory, streams, torch.distributed).
"""
from . import _lib  # noqa: F401

__all__ = ["api", "scheduler", "tilegen"]
__version__ = "0.1.0"
