"""Build the CUDA library IN-TREE for sm_100a: quasimodo_b200/libquasimodo_b200.so (C-ABI, no torch).

nvcc cross-compiles without a GPU, so this runs in the build container; the .so is git-ignored but
travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libquasimodo_b200.so")
DRIVER = os.path.join(HERE, "qm_driver")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h"))] + [os.path.join(HERE, "..", "include", "quasimodo_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_driver()
        return SO
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    subprocess.check_call([NVCC, "-shared", "-o", SO] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++"])
    build_driver(force=True)
    return SO


def build_driver(force=False):
    """the C++ host driver (csrc/driver/qm_driver.cpp) over the C-ABI: quasimodo_b200/qm_driver, rpath = its own directory"""
    src = os.path.join(CSRC, "driver", "qm_driver.cpp")
    hdr = os.path.join(HERE, "..", "include", "quasimodo_b200.h")
    if not force and os.path.exists(DRIVER) and os.path.getmtime(DRIVER) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(SO)):
        return DRIVER
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", DRIVER, src, "-L" + HERE, "-lquasimodo_b200", "-lz",
                           "-lpthread", "-Wl,-rpath,$ORIGIN", "-Wl,--allow-shlib-undefined"])
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
