#!/usr/bin/env python
"""Build another copy of the library with extra -D flags, for same-box A/B measurements (selected with QSB200_LIB):
    python tools/build_variant.py u4 -DQSB_FILL_U=4        ->  quicksand_b200/libqsb200_u4.so
"""
import concurrent.futures as cf, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quicksand_b200 import build as B

def main():
    tag, defs = sys.argv[1], sys.argv[2:]
    obj = os.path.join(B.HERE, "build", "variant_" + tag)
    os.makedirs(obj, exist_ok=True)
    digest = B.source_hash()
    def one(src):
        o = os.path.join(obj, src.replace(".cu", ".o"))
        extra = ['-DQSB_SOURCE_HASH="%s"' % digest] if src == "capi.cu" else []
        r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + defs + extra + ["-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
        open(o + ".log", "w").write(r.stdout + r.stderr)
        if r.returncode: sys.stderr.write(r.stdout + r.stderr); raise SystemExit("nvcc failed on " + src)
        return o
    with cf.ThreadPoolExecutor(max_workers=len(B.SOURCES)) as ex:
        objs = list(ex.map(one, B.SOURCES))
    lib = os.path.join(B.HERE, "libqsb200_%s.so" % tag)
    r = subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode: sys.stderr.write(r.stdout + r.stderr); raise SystemExit("link failed")
    print(lib)

if __name__ == "__main__":
    main()
