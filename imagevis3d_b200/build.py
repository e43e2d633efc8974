"""Builds libtvk_cuda.so (the C-ABI library of include/tvk_cuda.h) in-tree with nvcc
for sm_100a.  No JIT cache: the .so lives next to the sources so that it travels
to the GPU box with the repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtvk_cuda.so")

CU_SOURCES = ["tvk_context.cu", "tvk_raycast.cu", "tvk_visibility.cu", "tvk_builder.cu", "tvk_quantize.cu", "tvk_exchange.cu", "tvk_classic.cu", "tvk_rebrick.cu"]
CPP_SOURCES = ["host/tvk_host.cpp", "host/tvk_uvf.cpp", "host/tvk_rawio.cpp"]
# -fmad=false: multiply-adds are fused only where the source says fmaf(), so that the ray-marcher
# rounds exactly like the CPU oracle (see the floating-point contract in tvk_raycast.cu)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def sources():
    out = [os.path.join(CSRC, s) for s in CU_SOURCES]
    out += [os.path.join(CSRC, s) for s in CPP_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, "tvk_internal.h"), os.path.join(ROOT, "include", "tvk_cuda.h"),
                        os.path.join(ROOT, "include", "tvk_renderer.h"), os.path.join(ROOT, "include", "tvk_uvf.h"), os.path.join(CSRC, "host", "tvk_host.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out.decode())
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lz", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
