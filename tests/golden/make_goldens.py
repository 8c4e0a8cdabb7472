"""Extract the 26 log-density known-answer checks of the reference's own test suite
(/root/reference/test/testsuite.t:159-187, tolerance 1e-8 at testsuite.t:34) into
testsuite_logprob_goldens.json.  Run in the build container (the reference tree does not exist on
the GPU box); the JSON is committed.   python tests/golden/make_goldens.py
"""
import json
import math
import os
import re

SRC = "/root/reference/test/testsuite.t"
WHICH = {"bernoulli": 0, "uniform": 1, "gaussian": 2, "gamma": 3, "beta": 4, "binomial": 5, "poisson": 6,
         "categorical_array": 7, "categorical_vector": 7, "dirichlet_array": 8, "dirichlet_vector": 8}
VECS = {"mp": [0.2, 0.6, 0.2], "dv": [0.6, 0.3, 0.1], "dp": [1.0, 1.0, 1.0]}


def num(tok):
    tok = tok.strip()
    if tok in ("true", "false"):
        return 1.0 if tok == "true" else 0.0
    return float(tok)


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return [a.strip() for a in out]


def value(expr):
    expr = expr.strip()
    m = re.match(r"^(-?)M\.log\(([\d.]+)\)$", expr)
    if m:
        v = math.log(float(m.group(2)))
        return -v if m.group(1) else v
    return float(expr)


cases = []
for lineno, line in enumerate(open(SRC), 1):
    m = re.search(r'assertEq\("([^"]+)", \[D\.(\w+)(?:\(\d\))?\(double\)\]\.logprob\((.*)\), (.+)\)\s*$', line)
    if not m or not (159 <= lineno <= 187):
        continue
    name, dist, argstr, expected = m.groups()
    args = []
    for a in split_args(argstr):
        if a.startswith("array("):
            args += [float(t) for t in a[6:-1].split(",")]
        elif a in VECS:
            args += VECS[a]
        else:
            args.append(num(a))
    cases.append({"name": name, "dist": dist, "which": WHICH[dist], "args": args, "value": value(expected), "line": lineno})

assert len(cases) == 26, len(cases)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "testsuite_logprob_goldens.json")
json.dump({"source": "test/testsuite.t:159-187", "tolerance": 1e-8, "cases": cases}, open(out, "w"), indent=1)
print("wrote", out, len(cases))
