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

__all__ = ["abi", "install_jit_alias"]


def install_jit_alias():
    """Make `import jit` resolve to this package's extension (drop-in for the reference)."""
    mod = importlib.import_module("numpyllvm_b200.jit")
    sys.modules.setdefault("jit", mod)
    return mod
