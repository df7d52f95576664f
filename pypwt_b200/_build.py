"""In-tree build of the native pieces (no pip, no network):

    libpwt_b200.so   CUDA kernels + C ABI, nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
    pycudwt*.so      Cython wrapper over the C ABI, linked with rpath=$ORIGIN

Usage:  python pypwt_b200/_build.py        (or __graft_entry__.build())
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, cwd=cwd, check=True)


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library():
    _run(["make", "-j8", "-C", os.path.join(HERE, "csrc")])
    return os.path.join(HERE, "libpwt_b200.so")


def build_extension():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(HERE, "pycudwt" + ext)
    pyx = os.path.join(HERE, "pycudwt.pyx")
    hdr = os.path.join(ROOT, "include", "pwt_b200.h")
    cpp = os.path.join(HERE, "csrc", "build", "pycudwt.cpp")
    if _newer(target, [pyx, hdr]):
        os.makedirs(os.path.dirname(cpp), exist_ok=True)
        _run([sys.executable, "-m", "cython", "--cplus", "-3", "-o", cpp, pyx])
        _run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w",
              "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include"),
              cpp, "-o", target, "-L" + HERE, "-lpwt_b200", "-Wl,-rpath,$ORIGIN"])
    return target


def build_all():
    lib = build_library()
    ext = build_extension()
    return lib, ext


if __name__ == "__main__":
    print(build_all())
