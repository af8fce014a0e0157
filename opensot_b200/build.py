"""In-tree build of the native library (nvcc, sm_100a only). No JIT cache: the .so lands in
opensot_b200/lib/ so that it travels with the repo snapshot to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libopensot_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "opensot_b200.h"))
    deps.append(os.path.join(ROOT, "include", "opensot_b200_synth.h"))
    return deps


def build_native(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and not _newer(LIB, sources()):
        return LIB
    cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v" if verbose else "-warn-spills",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(CSRC, "osb_capi.cu"), os.path.join(CSRC, "osb_layout.c"),
           os.path.join(CSRC, "osb_synth.c"), os.path.join(CSRC, "sot_plan.cpp"),
           "-o", LIB, "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


def build_profile(block=100):
    """Development aid: the library with cycle stamps per phase of the register-row kernel
    (block `block` prints them; tests/gpu_prof.py runs it on the GPU box)."""
    out = os.path.join(LIBDIR, "libopensot_b200_prof.so")
    cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-DOSB_RR_PROFILE=%d" % block, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(CSRC, "osb_capi.cu"), os.path.join(CSRC, "osb_layout.c"),
           os.path.join(CSRC, "osb_synth.c"), os.path.join(CSRC, "sot_plan.cpp"), "-o", out, "-lcudart"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_profile() if "--profile" in sys.argv else build_native(force=True, verbose=True))
