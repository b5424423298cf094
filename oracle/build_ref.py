#!/usr/bin/env python3
"""Build the UNMODIFIED reference rasterizer as a checker: oracle/_ref/ref_rasterizer.so.

TEST INFRASTRUCTURE ONLY.  The product (4d-gaussian-splatting_b200/) never imports, links or
executes anything built here; only tests/, __graft_entry__.smoke() and bench.py's reference arm
do, and only to check / time the reference itself.

What it does: compiles the reference's five source files *where they lie* under
/root/reference (no copy into this repo), exactly the list the reference JIT-loads itself
(reference: gaussian_renderer/diff_gaussian_rasterization.py:22-27):

    diff-gaussian-rasterization/cuda_rasterizer/{rasterizer_impl,forward,backward}.cu
    diff-gaussian-rasterization/rasterize_points.cu
    diff-gaussian-rasterization/ext.cpp

with the reference's own include path (-I third_party/glm) plus one extra compiler flag,
`-include cstdint`: rasterizer_impl.h uses uint32_t / uintptr_t without including <cstdint>,
which gcc 13 rejects (flag only, no source edit).  The reference's own build systems
(setup.py / CMakeLists.txt / its JIT `load()`) are not run.  Target: sm_100a, so the reference
kernels run natively on the B200 next to ours.

Outputs go only to oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .so travels to the GPU
box; /root/reference itself does not exist there).
"""
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FDGS_REFERENCE_DIR", "/root/reference")
DGR = os.path.join(REF, "diff-gaussian-rasterization")
MODULE = "ref_rasterizer"

SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
    "ext.cpp",
]


def available():
    return all(os.path.exists(os.path.join(DGR, s)) for s in SOURCES)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))


def build(force=False, verbose=True):
    """Returns the path of the built module, or None when /root/reference is absent."""
    target = os.path.join(OUT, MODULE + ".so")
    if not available():
        return target if os.path.exists(target) else None
    if os.path.exists(target) and not force:
        newest = max(os.path.getmtime(os.path.join(DGR, s)) for s in SOURCES)
        if os.path.getmtime(target) > newest:
            return target
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include")]
    defs = ["-DTORCH_EXTENSION_NAME=" + MODULE, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    glm = os.path.join(DGR, "third_party", "glm")
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(DGR, s)
        obj = os.path.join(OUT, "obj", os.path.basename(s) + ".o")
        objs.append(obj)
        if s.endswith(".cu"):
            cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                   "-w", "-include", "cstdint", "-I", glm, "-I", DGR] + defs
            for i in inc:
                cmd += ["-isystem", i]
            cmd += ["-c", src, "-o", obj]
        else:
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-include", "cstdint", "-I", glm, "-I", DGR] + defs
            for i in inc:
                cmd += ["-isystem", i]
            cmd += ["-c", src, "-o", obj]
        jobs.append(cmd)
    if verbose:
        print("[oracle/_ref] compiling %d reference sources from %s" % (len(jobs), DGR), flush=True)
    with ThreadPoolExecutor(max_workers=5) as ex:
        list(ex.map(_run, jobs))
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    _run(["g++", "-shared", "-o", target] + objs +
         ["-L" + torch_lib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
          "-L" + os.path.join(cuda_home, "lib64"), "-lcudart", "-Wl,-rpath," + torch_lib])
    if verbose:
        print("[oracle/_ref] built", target, flush=True)
    return target


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference sources not present; nothing built")
