"""Build libdglb200.so (the C-ABI library of include/dglb200.h) for sm_100a, in-tree.

    python dgl-0.5-benchmark_b200/build.py [--force] [--verbose-ptxas]

Each csrc/*.cu is compiled to an object file in parallel and linked into lib/libdglb200.so.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libdglb200.so")
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose_ptxas=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "dglb200.h")]
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _newer(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-c", src, "-o", obj]
            jobs.append((src, cmd))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            futs = {ex.submit(subprocess.run, cmd, capture_output=True, text=True): src for src, cmd in jobs}
            for f in cf.as_completed(futs):
                r = f.result()
                if verbose_ptxas or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s" % futs[f])
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]
    if force or jobs or _newer(LIB, objs):
        # extern "C" entry points are exported explicitly (visibility=hidden elsewhere)
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    build_torch_extension(force=force or bool(jobs))
    return LIB


TORCH_LIB = os.path.join(HERE, "lib", "libdglb200_torch.so")


def build_torch_extension(force=False):
    """g++ -> lib/libdglb200_torch.so: the TORCH_LIBRARY("dglb200") ops of csrc_torch/ops.cpp, linked against
    lib/libdglb200.so (rpath $ORIGIN) and torch's own libraries.  No CUDA code is compiled here: the extension
    only needs torch's stream / allocator / device-guard headers."""
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(HERE, "csrc_torch", "ops.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "dglb200.h"), LIB]
    if not (force or _newer(TORCH_LIB, deps)):
        return TORCH_LIB
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", TORCH_LIB,
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), "-DUSE_CUDA",
           "-I" + os.path.join(cuda_home, "include")]
    cmd += ["-isystem" + p for p in ce.include_paths()]
    for lp in ce.library_paths():
        cmd += ["-L" + lp, "-Wl,-rpath," + lp]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-L" + os.path.dirname(LIB), "-ldglb200", "-Wl,-rpath,$ORIGIN", "-Wl,--no-as-needed"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the torch extension failed")
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose_ptxas="--verbose-ptxas" in sys.argv))
    print(TORCH_LIB)
