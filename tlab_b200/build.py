"""Build libtlab_gpu.so (sm_100a) in-tree with nvcc.  No JIT, no torch extension machinery:
the product is a plain C-ABI shared library."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtlab_gpu.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nccl_paths():
    """torch bundles NCCL 2.28 (headers + libnccl.so.2); the system has 2.27.  Prefer torch's."""
    cands = []
    try:
        import nvidia.nccl  # type: ignore
        base = list(nvidia.nccl.__path__)[0]
        cands.append((os.path.join(base, "include"), os.path.join(base, "lib")))
    except Exception:
        pass
    cands.append(("/usr/include", "/usr/lib/x86_64-linux-gnu"))
    for inc, lib in cands:
        if os.path.exists(os.path.join(inc, "nccl.h")):
            so = [f for f in glob.glob(os.path.join(lib, "libnccl.so*"))]
            if so:
                return inc, lib, sorted(so)[0]
    return None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + headers() + [os.path.join(HERE, "..", "include", "tlab_gpu.h"),
                                                               os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nccl = _nccl_paths()
    common = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp"] + ARCH
    if nccl:
        common += ["-I", nccl[0], "-DTLAB_HAVE_NCCL=1"]
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in headers())
                and os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "tlab_gpu.h"))):
            continue
        cmd = common + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            print(out)
    link = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-lcufft", "-Xcompiler", "-fopenmp"]
    if nccl:
        # link by full path so that the library loads next to torch's own NCCL on the GPU box
        link += ["-Xlinker", nccl[2], "-Xlinker", "-rpath", "-Xlinker", nccl[1]]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
