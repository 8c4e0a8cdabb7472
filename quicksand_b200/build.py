"""Build libqsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m quicksand_b200.build [--force]

Each translation unit is compiled to an object in parallel, then linked with ``nvcc -shared``.
The library links only the CUDA runtime: no torch, no Python -- Terra loads it directly.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqsb200.so")
SOURCES = ["capi.cu", "multi.cu", "glm_hmc.cu", "engine_f1.cu", "engine_f1w.cu", "engine_f2.cu", "engine_f4.cu", "engine_f5.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def source_hash():
    """SHA-256 over the CUDA sources and the public header, in a fixed order: compiled into the library (qsb_source_hash)
    and recomputed from the tree by bench.py, so a measured binary is tied to the sources it was built from."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "qsb200.h"))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    digest = source_hash()
    stamp = os.path.join(OBJ, "source_hash.txt")
    old = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if old != digest or not os.path.exists(LIB):
        # a different tree than the one the objects were built from: recompile what embeds the hash, relink
        try:
            os.remove(os.path.join(OBJ, "capi.o"))
        except OSError:
            pass
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "qsb200.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        extra = ['-DQSB_SOURCE_HASH="%s"' % digest] if os.path.basename(s) == "capi.cu" else []
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + extra + ["-c", s, "-o", o], capture_output=True, text=True)
        with open(o + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return s, r.returncode, r.stdout + r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for s, rc, out in ex.map(compile_one, jobs):
                if rc != 0:
                    sys.stderr.write(out)
                    raise RuntimeError("nvcc failed on %s" % s)
                if verbose:
                    print("compiled", os.path.basename(s))
    objs = [os.path.join(OBJ, src.replace(".cu", ".o")) for src in SOURCES]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        with open(stamp, "w") as f:
            f.write(digest + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
