"""In-tree build of the sm_100a extension (``mine_b200/ops/_mine_b200_cuda.so``).

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` for every ``.cu`` (cross-compiles
without a GPU), g++ for the binding files, one shared object next to this file so it travels with
the source snapshot to the GPU box.  Incremental: objects are rebuilt only when their source (or a
header) is newer.  ``python -m mine_b200.ops.build [--force] [--verbose]``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
NAME = "_mine_b200_cuda"
SO_PATH = os.path.join(HERE, NAME + ".so")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cutlass_include() -> List[str]:
    import importlib.util
    out = []
    for pkg, rel in (("flashinfer", "data/cutlass/include"), ("tilelang", "3rdparty/cutlass/include")):
        spec = importlib.util.find_spec(pkg)
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], rel)
            if os.path.isdir(p):
                out.append(p)
                break
    return out


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True)
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, lib, int(torch._C._GLIBCXX_USE_CXX11_ABI)


def _newer(src_files, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return r


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    inc, torch_lib, abi = _torch_paths()
    py_inc = sysconfig.get_paths()["include"]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    cutlass = _cutlass_include()
    jobs, objs = [], []
    for f in cu:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [nvcc] + NVCC_ARCH + ["-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                                        "-Xcompiler", "-fPIC", "-I" + CSRC, "-I" + os.path.join(cuda_home, "include")]
            cmd += ["-I" + c for c in cutlass]
            cmd += ["-c", src, "-o", obj]
            jobs.append(cmd)
    for f in cpp:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
                   f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-I" + CSRC, "-I" + py_inc, "-I" + os.path.join(cuda_home, "include")]
            cmd += ["-I" + p for p in inc]
            cmd += ["-c", src, "-o", obj]
            jobs.append(cmd)
    t0 = time.time()
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or _newer(objs, SO_PATH):
        link = ["g++", "-shared", "-o", SO_PATH] + objs + [
            "-L" + torch_lib, "-L" + os.path.join(cuda_home, "lib64"), "-Wl,-rpath," + torch_lib,
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
        _run(link, verbose)
    if verbose:
        print(f"built {SO_PATH} ({len(jobs)} objects, {time.time() - t0:.1f}s)")
    return SO_PATH


def load():
    """Import the extension module, building it if the shared object is missing."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    if not os.path.exists(SO_PATH):
        build()
    spec = importlib.util.spec_from_file_location(NAME, SO_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(SO_PATH)
