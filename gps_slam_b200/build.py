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
    ("tsdf_mesh.cu", ["-fmad=false"]),
    ("tsdf_engine.cu", ["-fmad=false"]),
    ("icp_kernels.cu", ["-fmad=false"]),
    ("gs_project.cu", ["-fmad=false"]),
    ("gs_raster.cu", []),
    ("gs_spawn.cu", ["-fmad=false"]),
    ("gs_staged.cu", ["-fmad=false"]),
    ("gs_raw.cu", []),
    ("gs_ssim.cu", []),
    ("gs_knn.cu", []),
    ("gs_comm.cu", []),
    ("peer_mbox.cu", []),
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


# ---- C++ host layer over the C ABI (cxx/): the reference-facing gsplat interface (gsplat_wapper.hpp classes) compiled against the
# torch wheel's libtorch, registered as torch ops for the tests; and the InfiniTAM-facing facade (header-only, built into a test driver)
CXX_DIR = os.path.join(HERE, "cxx")
TORCH_SHIM = os.path.join(HERE, "libgsplat_b200_torch.so")
CXX_SOURCES = [os.path.join("gsplat", "gsplat_b200.cpp"), "gsplat_b200_ops.cpp"]


def _newer_than(target, paths):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    for p in paths:
        if os.path.isdir(p):
            for root, _, files in os.walk(p):
                if any(os.path.getmtime(os.path.join(root, f)) > t for f in files):
                    return True
        elif os.path.getmtime(p) > t:
            return True
    return False


def build_torch_shim(force=False):
    """libgsplat_b200_torch.so: cxx/gsplat/*.{hpp,cpp} + the op registration, g++ against libtorch (2 translation units, ~2 min)."""
    inc = os.path.join(HERE, "..", "include")
    if not force and not _newer_than(TORCH_SHIM, [os.path.join(CXX_DIR, "gsplat"), os.path.join(CXX_DIR, "gsplat_b200_ops.cpp"),
                                                  os.path.join(inc, "gpsslam_b200.h")]):
        return TORCH_SHIM
    import torch
    tdir = os.path.dirname(torch.__file__)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-std=c++17", "-O2", "-fPIC", "-w", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
             "-I", CXX_DIR, "-I", os.path.join(CXX_DIR, "gsplat"), "-I", inc, "-I", os.path.join(tdir, "include"),
             "-I", os.path.join(tdir, "include", "torch", "csrc", "api", "include"), "-I", "/usr/local/cuda/include"]
    procs, objs = [], []
    for src in CXX_SOURCES:
        obj = os.path.join(objdir, os.path.basename(src).replace(".cpp", ".o"))
        procs.append((src, subprocess.Popen(["g++"] + flags + ["-c", os.path.join(CXX_DIR, src), "-o", obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise RuntimeError("g++ failed on %s" % src)
    tlib = os.path.join(tdir, "lib")
    subprocess.check_call(["g++", "-shared", "-o", TORCH_SHIM] + objs + [
        "-L", HERE, "-lgpsslam_b200", "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
        "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib])
    return TORCH_SHIM


ITM_DRIVER = os.path.join(HERE, "build", "itm_facade_driver")
ITM_DRIVER_SRC = os.path.join(HERE, "..", "tests", "cxx", "itm_facade_driver.cpp")


def itm_facade_flags():
    """compile + link flags of a program that uses the InfiniTAM-facing facade (cxx/InfiniTAM) -- also used by oracle/itm_ref/Makefile"""
    inc = ["-I", os.path.join(CXX_DIR, "InfiniTAM"), "-I", os.path.join(HERE, "..", "include"), "-I", "/usr/local/cuda/include"]
    link = ["-L", HERE, "-lgpsslam_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + HERE]
    return inc, link


def build_itm_driver(force=False):
    """tests/cxx/itm_facade_driver.cpp against the facade headers: the C++ route into the TSDF / ICP engine that tests/test_cxx_itm_gpu.py runs"""
    if not force and not _newer_than(ITM_DRIVER, [ITM_DRIVER_SRC, os.path.join(CXX_DIR, "InfiniTAM"), LIB]):
        return ITM_DRIVER
    os.makedirs(os.path.dirname(ITM_DRIVER), exist_ok=True)
    inc, link = itm_facade_flags()
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall"] + inc + [ITM_DRIVER_SRC, "-o", ITM_DRIVER] + link)
    return ITM_DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--cxx" in sys.argv:
        print(build_torch_shim(force="--force" in sys.argv))
        print(build_itm_driver(force="--force" in sys.argv))
