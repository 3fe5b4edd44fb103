"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/b200conv.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as ge

ROOT = ge.ROOT


@pytest.fixture(scope="module")
def pkg():
    p = ge.load()
    p.build()
    return p


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200conv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200conv_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_convolver_surface():
    names = declared_symbols()
    for need in ("b200conv_create", "b200conv_free", "b200conv_init", "b200conv_init_range",
                 "b200conv_destroy", "b200conv_process", "b200conv_process_device",
                 "b200conv_data_size", "b200conv_rank", "b200conv_fastconv_parse",
                 "b200conv_fastconv_apply", "b200conv_fastconv_parse_apply",
                 "b200conv_fastconv_restore"):
        assert need in names


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    # and the Python mirror binds exactly that set
    assert sorted(pkg._SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200ConvError) as e:
        pkg.ConvolverBatch(1, 0)
    assert e.value.code == pkg.ERR_CUDA
    # queries on a NULL handle are harmless
    assert pkg.lib().b200conv_rank(None, 0) == 0
    assert pkg.lib().b200conv_version().startswith(b"b200conv")


def test_library_is_sm100a_sass(pkg):
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    for dirpath, _, files in os.walk(ge.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


@pytest.mark.parametrize("in_len,inv_len,limit", [
    ([1000], 300, 256), ([5000, 3000, 5000], 4096, 1024), ([40000, 100], 9000, 0),
    ([70000], 70000, 32768), ([128, 129, 127], 128, 128), ([10], 3, 1), ([2000], 2500, 100000),
])
def test_chirp_plan_matches_the_reference_arithmetic(pkg, in_len, inv_len, limit):
    """b200conv_chirp_plan is host arithmetic (SyncChirpProcessor.cpp:1224-1250,1299-1331): checked
    here against the CPU restatement and the independent model, no GPU needed."""
    import chirp_model
    from oracle import bindings
    pkg.lib()
    got = pkg.chirp_plan(in_len, inv_len, limit)
    assert got == bindings.chirp_plan(in_len, inv_len, limit)
    assert got == chirp_model.plan(in_len, inv_len, limit)
