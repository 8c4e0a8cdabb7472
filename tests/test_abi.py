"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/qsb200.h declares
(no compute calls: there is no GPU here), the product has no route into the oracle, and the host-side
mirror of the reference interface keeps the reference's names, defaults and error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "qsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qsb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(qs):
    from quicksand_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libqsb200.so does not export %s" % name
    assert sorted(_lib.PROTOTYPES) == declared
    assert lib.qsb_version() == 100


def test_header_is_plain_c(tmp_path):
    """terralib.includec needs a header a C compiler accepts."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "qsb200.h"\nint main(void){ qsb_run_config c; c.seed = 1; return (int)c.seed - 1 + (QSB_VERSION != 100); }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/."""
    pkg = os.path.join(ROOT, "quicksand_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".t")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), (dirpath, f)
                assert "libqsoracle" not in text and "oracle/" not in text, (dirpath, f)


def test_kernel_constructor_defaults_match_reference(qs):
    from quicksand_b200 import _lib as L
    h = qs.HMCKernel().specs[0]        # hmc.t:57-66
    assert (h.kind, h.stepSize, h.numSteps, h.doAdapt) == (L.KERNEL_HMC, 1.0, 1, 1)
    h = qs.HMCKernel({"numSteps": 10, "doStepSizeAdapt": False}).specs[0]
    assert (h.numSteps, h.doAdapt) == (10, 0)
    d = qs.DriftKernel().specs[0]      # drift.t:63-73
    assert (d.kind, d.scale, d.doAdapt) == (L.KERNEL_DRIFT, 1.0, 1)
    r = qs.HARMKernel(scale=2.0).specs[0]   # harm.t:16-22
    assert (r.kind, r.scale, r.doAdapt) == (L.KERNEL_HARM, 2.0, 1)
    m = qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16)], [0.5, 0.5])
    assert [s.kind for s in m.specs] == [L.KERNEL_HARM, L.KERNEL_HMC] and [s.weight for s in m.specs] == [0.5, 0.5]
    with pytest.raises(AssertionError):   # mcmc.t:200-201
        qs.MixtureKernel([qs.HARMKernel()], [0.5, 0.5])
    with pytest.raises(TypeError):
        qs.HMCKernel(stepsize=1.0)


def test_no_gpu_means_loud_failure(qs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quicksand_b200 import _lib
    with pytest.raises(_lib.QsbError):
        qs.programs.funnel(8).model()


def test_cxx_mirror_compiles_links_and_fails_loudly_without_gpu(qs, tmp_path):
    """include/qsb200.hpp (header-only C++ mirror of the reference interface) against libqsb200.so: the reference's
    10-gaussian HMC test (testsuite.t:719-730).  Without a GPU the program must exit with the library's error, not fall back."""
    import subprocess
    import torch
    from quicksand_b200 import _lib
    exe = tmp_path / "example"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx", "example.cpp"),
                    "-o", str(exe), "-L", libdir, "-lqsb200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "qsb200:" in r.stderr


def test_indep_descriptor_layout(qs):
    """Host logic of qs.programs.indep: entries -> per-component descriptor arrays (vector choices expand to K components,
    gammamv / betamv reparametrise, latent parameters refer to COMPONENT indices and must point backwards), and the ctypes
    mirror of qsb_model_desc has the header's field order."""
    from quicksand_b200 import _lib as L
    lat = qs.programs.latent
    p = qs.programs.indep([("gamma", 3.0, 1.0), ("dirichlet", [2.0, 3.0, 5.0]), ("gamma", lat(0), 2.0), ("gammamv", 4.0, 8.0),
                           ("gaussian", lat(3, a=0.5, b=2.0), lat(2, a=0.1, exp=True))], ret="vector", lexids=[0, 1, 0, 2, 3],
                          factors=[None, None, (1.0, 0.5), None, None])
    a = p.arrays
    assert p.dim == 7 and p.retdim == 7
    assert list(a["erp_kind"]) == [L.ERP_GAMMA] + [L.ERP_DIRICHLET] * 3 + [L.ERP_GAMMA, L.ERP_GAMMA, L.ERP_GAUSSIAN]
    assert list(a["erp_p1"][1:4]) == [0.0, 1.0, 2.0]                       # index inside the vector choice
    assert a["erp_p0"][5] == 4.0 * 4.0 / 8.0 and a["erp_p1"][5] == 8.0 / 4.0   # gammamv(m, v) -> gamma(m*m/v, v/m)
    assert list(a["erp_lexid"]) == [0, 1, 1, 1, 0, 2, 3]
    assert list(a["erp_pref"]) == [-1, -1, -1, -1, -1, -1, -1, -1, 0, -1, -1, -1, 5, 4]   # entry 3 -> component 5, entry 2 -> component 4
    assert list(a["erp_ptf"]) == [0] * 13 + [1] and a["erp_pcoef"][12] == 2.0 and a["erp_p0"][6] == 0.5 and a["erp_p1"][6] == 0.1
    assert a["erp_fsigma"][4] == 0.5 and a["erp_fmu"][4] == 1.0 and a["erp_fsigma"][0] == 0.0
    with pytest.raises(AssertionError):
        qs.programs.indep([("gaussian", lat(0), 1.0)])                      # a parameter cannot read its own (or a later) choice
    with pytest.raises(AssertionError):
        qs.programs.indep([("gaussian", 0.0, 1.0), ("uniform", lat(0), 2.0)])   # a uniform's parameters are its bounds
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "qsb200.h")).read()
    body = hdr[hdr.index("typedef struct {", hdr.index("All pointers are HOST pointers")):hdr.index("} qsb_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)\s*;", body)
    assert names == [f[0] for f in L.ModelDesc._fields_], (names, [f[0] for f in L.ModelDesc._fields_])
