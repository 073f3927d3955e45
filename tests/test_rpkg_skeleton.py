"""The R package skeleton (rpkg/) cannot be installed here (no R in the image); these checks keep it
consistent with the library it wraps: the Makevars compiles exactly the CUDA sources build.py compiles, with the same
sm_100a flags; the .Call stubs name routines the shim registers, with matching arities; shim.c compiles against the
stand-in R headers."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RPKG = os.path.join(ROOT, "rpkg")


def test_makevars_compiles_the_same_sources_with_the_same_arch():
    from gpirt_b200 import build
    mk = open(os.path.join(RPKG, "src", "Makevars")).read()
    srcs = re.search(r"^CU_SOURCES\s*=\s*(.*)$", mk, re.M).group(1).split()
    assert srcs == build.SOURCES
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk
    for s in srcs:
        assert os.path.exists(os.path.join(ROOT, "gpirt_b200", "csrc", s))


def test_call_stubs_match_the_registered_routines():
    shim = open(os.path.join(ROOT, "gpirt_b200", "csrc", "rshim", "gpirt_rshim.c")).read()
    registered = dict((name, int(arity)) for name, arity in re.findall(r'\{"(_gpirt_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', shim))
    assert registered == {"_gpirt_gpirtMCMC": 7, "_gpirt_gpirtMCMC_b200": 10}
    for f in ("RcppExports.R", "gpirtMCMC_b200.R"):
        src = open(os.path.join(RPKG, "R", f)).read()
        for mt in re.finditer(r"\.Call\(`(_gpirt_\w+)`,", src):
            depth, nargs, i = 1, 1, mt.end()
            while depth:                       # count the top-level commas up to the matching parenthesis
                c = src[i]
                depth += c == "("
                depth -= c == ")"
                nargs += c == "," and depth == 1
                i += 1
            assert registered[mt.group(1)] == nargs, (f, mt.group(1), nargs)
    ns = open(os.path.join(RPKG, "NAMESPACE")).read()
    directives = [ln for ln in ns.splitlines() if ln.strip() and not ln.startswith("#")]
    assert "useDynLib(gpirt, .registration = TRUE)" in directives and not any("Rcpp" in ln for ln in directives)


def test_shim_translation_unit_of_the_package_compiles(tmp_path):
    out = tmp_path / "shim.o"
    subprocess.check_call(["gcc", "-c", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "fake_r", "include"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(RPKG, "src", "shim.c"), "-o", str(out)],
                          cwd=os.path.join(RPKG, "src"))
    assert out.exists()
