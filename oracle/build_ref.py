"""TEST INFRASTRUCTURE ONLY -- builds the reference's own CUDA op into oracle/_ref/.

The reference extension (`groundingdino._C`, reference setup.py:78-83, csrc/vision.cpp:53-56) is
compiled from its sources *in place* under /root/reference; nothing is copied into this repo and
the only outputs are object files and `ref_C.so` under the git-ignored `oracle/_ref/`.  The single
incompatibility with torch 2.11 (csrc/MsDeformAttn/ms_deform_attn_cuda.cu:65,:135) is bridged by
force-including `oracle/ref_shim.h`; the reference's own build system (setup.py) is not run.

The resulting `.so` is a torch extension for sm_100a.  It is only ever *loaded* by the `-m gpu`
tests and by `bench.py (key `ref_cuda_us_per_layer`)` as the GPU-side A/B oracle; the product never imports it.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/groundingdino/models/GroundingDINO/csrc"
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "ref_C.so")


def ref_so_path():
    return SO


def build(verbose=False):
    """Compile the reference op. Returns the .so path, or None when /root/reference is absent."""
    if not os.path.isdir(REF_CSRC):
        return SO if os.path.exists(SO) else None
    srcs = [
        os.path.join(REF_CSRC, "vision.cpp"),
        os.path.join(REF_CSRC, "MsDeformAttn", "ms_deform_attn_cpu.cpp"),
        os.path.join(REF_CSRC, "MsDeformAttn", "ms_deform_attn_cuda.cu"),
        os.path.join(REF_CSRC, "cuda_version.cu"),
    ]
    shim = os.path.join(HERE, "ref_shim.h")
    newest = max(os.path.getmtime(p) for p in srcs + [shim, __file__])
    if os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    os.makedirs(OUT, exist_ok=True)

    import torch
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths(device_type="cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"], "-I", REF_CSRC]
    common = [
        "-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=ref_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-include", shim, "-O2", "-std=c++17", "-w",
    ]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    objs = []
    for s in srcs:
        o = os.path.join(OUT, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", s, "-o", o, "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC", abi,
                   "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                   "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                   "--expt-relaxed-constexpr"] + common + inc
        else:
            cmd = ["g++", "-c", s, "-o", o, "-fPIC", abi] + common + inc
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    libdirs = ce.library_paths(device_type="cuda")
    link = ["g++", "-shared", "-o", SO] + objs
    for d in libdirs:
        link += ["-L", d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
             "-lcudart"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    return SO


def load():
    """Import the built reference op (needs a GPU to *run*, not to import). None if not built."""
    if not os.path.exists(SO):
        return None
    import importlib.util
    import torch  # noqa: F401  (must be imported first so libtorch symbols resolve)

    spec = importlib.util.spec_from_file_location("ref_C", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference op:", p)
