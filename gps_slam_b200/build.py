"""Builds gps_slam_b200/libgpsslam_b200.so (the C-ABI engine) with nvcc for sm_100a, in-tree.

TSDF / ICP / pose sources are compiled with -fmad=false (device) and -ffp-contract=off (host) so that the fp32
results are bit-identical to the reference's shared math (see csrc/tsdf_kernels.cu header)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpsslam_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
          "-I", os.path.join(HERE, "..", "include")]
# (source, extra flags)
SOURCES = [
    ("tsdf_kernels.cu", ["-fmad=false"]),
    ("tsdf_engine.cu", ["-fmad=false"]),
    ("icp_kernels.cu", ["-fmad=false"]),
    ("gs_project.cu", ["-fmad=false"]),
    ("gs_raster.cu", []),
    ("gs_spawn.cu", ["-fmad=false"]),
    ("gs_staged.cu", ["-fmad=false"]),
    ("gs_raw.cu", []),
    ("gs_ssim.cu", []),
    ("gs_engine.cu", ["-fmad=false"]),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return os.path.getmtime(os.path.join(HERE, "..", "include", "gpsslam_b200.h")) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
