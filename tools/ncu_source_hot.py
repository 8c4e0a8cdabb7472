#!/usr/bin/env python
"""Hot regions of a kernel from the `ncu --page source --csv` export (SASS view): basic blocks (split at branch targets are not
known here, so: split at BRA/CALL/RET/EXIT/BSYNC) with their share of the sampled stalls and of the executed instructions.
    python tools/ncu_source_hot.py gpurun_out/x_source.csv.gz [min_pct]
"""
import csv, gzip, sys, io

def main():
    path = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    S = ix["# Samples"]; E = ix["Instructions Executed"]; T = ix["Thread Instructions Executed"]
    stall_cols = [(h, i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(int(r[S] or 0) for r in data); tot_e = sum(int(r[E] or 0) for r in data)
    print("instructions %d, samples %d, warp-instructions executed %d" % (len(data), tot_s, tot_e))
    blocks = []; cur = []
    for k, r in enumerate(data):
        cur.append(k)
        op = r[ix["Source"]].split()
        opn = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
        if opn.split(".")[0] in ("BRA", "CALL", "RET", "EXIT", "BSYNC", "BREAK", "WARPSYNC"):
            blocks.append(cur); cur = []
    if cur: blocks.append(cur)
    out = []
    for b in blocks:
        s = sum(int(data[k][S] or 0) for k in b); e = sum(int(data[k][E] or 0) for k in b)
        out.append((s, e, b))
    print("%8s %6s %6s %6s  %s" % ("first", "n", "samp%", "exec%", "top stalls / opcode mix"))
    for s, e, b in out:
        if 100.0 * s / max(tot_s, 1) < minpct and 100.0 * e / max(tot_e, 1) < minpct: continue
        st = {}
        for h, i in stall_cols:
            v = sum(int(data[k][i] or 0) for k in b)
            if v: st[h[6:]] = v
        top = sorted(st.items(), key=lambda x: -x[1])[:4]
        ops = {}
        for k in b:
            op = data[k][ix["Source"]].split()
            opn = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
            opn = opn.split(".")[0]
            ops[opn] = ops.get(opn, 0) + 1
        mix = sorted(ops.items(), key=lambda x: -x[1])[:6]
        print("%8d %6d %6.2f %6.2f  %s | %s" % (b[0], len(b), 100.0 * s / max(tot_s, 1), 100.0 * e / max(tot_e, 1),
              " ".join("%s:%d" % (a, 100 * v // max(s, 1)) for a, v in top), " ".join("%s%d" % m for m in mix)))

if __name__ == "__main__":
    main()
