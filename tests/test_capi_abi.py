"""The C-ABI library loads on a box without a GPU, exports every symbol include/gpirt_b200.h declares, and refuses to
compute without a CUDA device (no CPU fallback) — no compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpirt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpirt_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_loader_agree():
    from gpirt_b200 import _lib
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from gpirt_b200 import _lib
    L = _lib.load()
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_no_cpu_fallback():
    from gpirt_b200 import _lib
    import gpirt_b200.sampler as G
    L = _lib.load()
    if L.gpirt_b200_device_count() > 0:
        pytest.skip("a GPU is present; the refusal path is only observable without one")
    with pytest.raises(_lib.GpirtError) as e:
        G.se_cov(np.zeros(3), np.zeros(3))
    assert e.value.status == _lib.ERR_CUDA
    with pytest.raises(_lib.GpirtError):
        G.Sampler(np.array([[1.0, -1.0], [-1.0, 1.0]]), np.zeros(2))
    assert L.gpirt_b200_strerror(_lib.ERR_NOT_PD).decode() == "chol(): decomposition failed"


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gpirt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle's", "").replace("oracle/gpo_rng.h", "") or f in ("philox.cuh",), \
                    "%s mentions the oracle" % os.path.join(dirpath, f)


def test_opts_struct_layout_matches_header(tmp_path):
    """the ctypes mirror of gpirt_b200_opts against offsetof() of the real header, compiled with the C compiler"""
    import subprocess
    from gpirt_b200._lib import Opts
    fields = [name for name, _ in Opts._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gpirt_b200.h"\nint main(void) {\n'
                   + "".join('  printf("%s %%zu\\n", offsetof(gpirt_b200_opts, %s));\n' % (f, f) for f in fields)
                   + '  printf("sizeof %zu\\n", sizeof(gpirt_b200_opts));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for f in fields:
        assert int(got[f]) == getattr(Opts, f).offset, f
    assert int(got["sizeof"]) == C.sizeof(Opts)
    assert Opts.seed.offset == 0 and Opts.rank.offset == 24 and Opts.nccl_unique_id.offset == 48   # round-1 prefix unchanged
