"""In-tree build of the native parts (sm_100a only).

  libnlkernels.so          csrc/nl_plan.cpp + nl_kernels.cu + nl_runtime.cu   (nvcc, static cudart)
  jit.cpython-312-*.so     csrc/jit/*.cpp — the `jit` Python extension        (g++)

Everything is compiled with explicit commands (no JIT cache): the built .so
files live next to this file so they travel to the GPU box with the snapshot.
`-fmad=false` keeps float multiply/add chains un-contracted, which the
bit-exact parity contract relies on.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SUFFIX = os.environ.get("NL_BUILD_SUFFIX", "")     # A/B builds: build<suffix>/ and libnlkernels<suffix>.so (use with NL_EXTRA_NVCC_FLAGS)
BUILD = os.path.join(ROOT, "build" + SUFFIX)
INCLUDE = os.path.join(ROOT, "include")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", f"-I{INCLUDE}", f"-I{CSRC}"]
NVCC_FLAGS += os.environ.get("NL_EXTRA_NVCC_FLAGS", "").split()   # tuning experiments (-DNL_PIPE_ROWS=8 ...); use with --force

KERNEL_SOURCES = ["nl_plan.cpp", "nl_kernels.cu", "nl_runtime.cu", "nl_sort.cu", "nl_checksum.cu", "nl_cast.cu",
                  "nl_dynamic_i8.cu", "nl_dynamic_i16.cu", "nl_dynamic_u8.cu", "nl_dynamic_u16.cu", "nl_dynamic_u32.cu", "nl_dynamic_u64.cu",
                  "nl_dynamic_i32.cu", "nl_dynamic_i64.cu", "nl_dynamic_f32.cu", "nl_dynamic_f64.cu",
                  "nl_static_i32.cu", "nl_static_i64.cu", "nl_static_f32.cu", "nl_static_f64.cu",
                  "nl_static_i32_b.cu", "nl_static_i64_b.cu", "nl_static_f32_b.cu", "nl_static_f64_b.cu",
                  "nl_family_i32.cu", "nl_family_i64.cu", "nl_family_f32.cu", "nl_family_f64.cu",
                  "nl_family_i32_b.cu", "nl_family_i64_b.cu", "nl_family_f32_b.cu", "nl_family_f64_b.cu"]
JIT_SOURCES = ["jitpackage.cpp", "device_sort.cpp", "thunk.cpp", "operation.cpp", "thunk_as_number.cpp", "thunk_as_mapping.cpp",
               "thunk_as_sequence.cpp", "thunk_methods.cpp", "parser.cpp", "optimizer.cpp", "scheduler.cpp",
               "debug_printer.cpp", "storage.cpp"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"build command failed ({res.returncode}): {' '.join(cmd)}")
    return res.stdout


def _headers(d):
    return [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".hpp", ".inc", ".cuh"))]


def build_kernels(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    """nvcc -> numpyllvm_b200/libnlkernels.so"""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(PKG, "libnlkernels" + SUFFIX + ".so")
    hdrs = _headers(CSRC) + _headers(INCLUDE) + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src in KERNEL_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-x", "cu", "-c", s, "-o", o]
            jobs.append(cmd)
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or _stale(out, objs):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-ldl", "-lpthread"], verbose)
        import ctypes
        ctypes.CDLL(out)      # an undefined symbol shows up here, not on the GPU box (no device is needed to load)
    if ptxas_info:
        print("\n".join(logs))
    return out


def build_jit(force: bool = False, verbose: bool = False) -> str:
    """g++ -> numpyllvm_b200/jit.cpython-312-x86_64-linux-gnu.so (links libnlkernels via $ORIGIN rpath)"""
    import numpy
    jdir = os.path.join(CSRC, "jit")
    srcs = [os.path.join(jdir, s) for s in JIT_SOURCES if os.path.exists(os.path.join(jdir, s))]
    if not srcs:
        return ""
    os.makedirs(os.path.join(BUILD, "jit"), exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(PKG, "jit" + ext)
    hdrs = _headers(jdir) + _headers(INCLUDE) + [os.path.abspath(__file__)]
    flags = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function",
             f"-I{INCLUDE}", f"-I{jdir}", f"-I{sysconfig.get_paths()['include']}", f"-I{numpy.get_include()}",
             "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(BUILD, "jit", os.path.basename(s)[:-4] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([CXX] + flags + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or _stale(out, objs):
        _run([CXX, "-shared", "-o", out] + objs + [f"-L{PKG}", "-lnlkernels", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"], verbose)
    return out


def build_all(force: bool = False, verbose: bool = False):
    k = build_kernels(force=force, verbose=verbose)
    j = build_jit(force=force, verbose=verbose)
    return k, j


if __name__ == "__main__":
    force = "--force" in sys.argv
    k = build_kernels(force=force, verbose=True, ptxas_info="--ptxas" in sys.argv)
    j = build_jit(force=force, verbose=True)
    print("built", k, j)
