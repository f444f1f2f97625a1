"""Block layouts analysed once (``bsls_plan``) and cached per ``blocks`` tensor.

The reference re-narrows ``blocks`` to C ints on every call
(python/c_extensions/c_extensions.pyx:36-38); on the GPU the layout analysis (validation,
uniform-size detection, tile / large-block binning) is done once and reused.
"""
import collections
import ctypes

import numpy as np
import torch

from . import _lib


class BlockPlan:
    """Owns one ``bsls_plan`` handle.  ``blocks`` are start offsets (reference convention)."""

    def __init__(self, blocks, n, device=None):
        L = _lib.lib()
        if isinstance(blocks, BlockPlan):
            raise TypeError("already a plan")
        if not torch.is_tensor(blocks):
            blocks = torch.as_tensor(np.ascontiguousarray(blocks, dtype=np.int64))
        assert blocks.dim() == 1 and blocks.numel() > 0, "blocks must be a non-empty 1-D array"
        if device is None:
            device = blocks.device if blocks.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.n = int(n)
        self.numblocks = int(blocks.numel())
        assert self.n < 2 ** 31, "int32 indices as in the reference's C layer (proj_simplex.h:37)"
        b32 = blocks.to(device=self.device, dtype=torch.int32).contiguous()
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(L.bsls_plan_create(b32.data_ptr(), self.numblocks, self.n, stream, ctypes.byref(self._handle)),
                       "plan_create")
        info = (ctypes.c_int64 * 8)()
        _lib.check(L.bsls_plan_info(self._handle, ctypes.byref(info)))
        self.first, self.uniform, self.min_size, self.max_size = int(info[2]), int(info[3]), int(info[4]), int(info[5])
        self.tiles, self.large = int(info[6]), int(info[7])

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().bsls_plan_destroy(h)
            except Exception:
                pass
            self._handle = None


_CACHE = collections.OrderedDict()
_CACHE_MAX = 32


def plan_for(blocks, n, device):
    """Cached plan for a ``blocks`` tensor (keyed on identity + version counter; the cache
    keeps the tensor alive so its address cannot be recycled under us)."""
    if isinstance(blocks, BlockPlan):
        assert blocks.n == n, "plan was built for n=%d, got %d" % (blocks.n, n)
        return blocks
    if not torch.is_tensor(blocks):
        return BlockPlan(blocks, n, device)
    # a plan's scratch serves one stream at a time (include/bsls_b200.h): one plan per (layout, stream)
    stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
    key = (id(blocks), blocks._version, int(n), str(device), stream)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is blocks:
        _CACHE.move_to_end(key)
        return hit[1]
    plan = BlockPlan(blocks, n, device)
    _CACHE[key] = (blocks, plan)
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)
    return plan


def clear_cache():
    _CACHE.clear()
