"""numpyllvm_b200 — B200-native execution engine for numpyllvm's lazy thunk pipelines.

Layout (only what the hot path needs):
  csrc/          CUDA kernels (sm_100a) + runtime + the C-ABI  -> libnlkernels.so
  csrc/jit/      the `jit` CPython extension: thunk type, parser, plan optimiser,
                 shard scheduler (mirrors the reference's Python surface)
  abi.py         ctypes mirror of include/nl_abi.h
  dist.py        one-process-per-GPU plumbing (torch.distributed) for bench / tests

`import numpyllvm_b200` does not load CUDA; `numpyllvm_b200.jit` is the drop-in
for the reference's `import jit`.
"""
import importlib
import sys

__all__ = ["abi", "install_jit_alias", "device_shards", "DeviceShard"]


def install_jit_alias():
    """Make `import jit` resolve to this package's extension (drop-in for the reference)."""
    mod = importlib.import_module("numpyllvm_b200.jit")
    sys.modules.setdefault("jit", mod)
    return mod


class DeviceShard:
    """One shard of an evaluated thunk where it lives: exposes ``__cuda_array_interface__`` (v3), so
    ``torch.as_tensor(shard, device=f"cuda:{shard.device}")`` / ``cupy.asarray(shard)`` see the data
    without a copy.  Keeps the thunk alive; the memory is the thunk's own storage (do not assign to
    the thunk while a view is in use)."""

    def __init__(self, owner, index, device, ptr, start, count, typestr):
        self.owner, self.index, self.device, self.start, self.count = owner, index, device, start, count
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}

    def __len__(self):
        return self.count

    # DLPack (torch.from_dlpack(shard), cupy.from_dlpack(shard), jax.dlpack.from_dlpack(shard))
    def __dlpack__(self, stream=None, **_):
        return self.owner._dlpack(self.index)      # the data is complete: the export synchronises first

    def __dlpack_device__(self):
        return (2, self.device)                    # (kDLCUDA, CUDA ordinal)


def device_shards(thunk):
    """Evaluate `thunk` and return its shards as DeviceShard views, in global index order
    (the host<->device edge of SURVEY 8f-3: results can stay on the GPUs)."""
    return [DeviceShard(thunk, i, *t) for i, t in enumerate(thunk.device_shards())]
