"""The C++ drop-in facade lsp::dspu::Convolver (lsp-dsp-units_b200/host) above the C ABI."""
import os
import subprocess

import pytest

import __graft_entry__ as ge

HOST = os.path.join(ge.PKG_DIR, "host")
REF_INC = "/root/reference/include"


def _make():
    ge.load().build()
    out = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_facade_builds_and_has_the_reference_public_surface():
    _make()
    hdr = open(os.path.join(HOST, "include", "lsp-plug.in", "dsp-units", "util", "Convolver.h")).read()
    for decl in ("bool init(const float *data, size_t count, size_t rank, float phase);",
                 "void process(float *dst, const float *src, size_t count);",
                 "size_t data_size() const;", "size_t rank() const;",
                 "void dump(IStateDumper *v) const;", "void construct();", "void destroy();",
                 "Convolver(const Convolver &) = delete;", "Convolver(Convolver &&) = delete;"):
        assert decl in hdr, decl
    syms = subprocess.run(["nm", "-DC", os.path.join(HOST, "libb200conv_host.so")],
                          capture_output=True, text=True).stdout
    for name in ("lsp::dspu::Convolver::init(float const*, unsigned long, unsigned long, float)",
                 "lsp::dspu::Convolver::process(float*, float const*, unsigned long)",
                 "lsp::dspu::Convolver::destroy()", "lsp::dspu::Convolver::dump("):
        assert name in syms, name


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="needs the reference headers")
def test_facade_dump_compiles_against_the_reference_state_dumper(tmp_path):
    """dump() against the reference's own IStateDumper.h (lsp-common-lib types come from the
    oracle shim, as that library is not available offline)."""
    out = subprocess.run(
        ["g++", "-std=c++11", "-fsyntax-only", "-DB200CONV_WITH_STATE_DUMPER", "-DLSP_DSP_UNITS_BUILTIN",
         "-I", os.path.join(HOST, "include"), "-I", os.path.join(ge.ROOT, "include"),
         "-I", REF_INC, "-I", os.path.join(ge.ROOT, "oracle", "shim"),
         os.path.join(HOST, "Convolver.cpp")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def _build_dump_probe(tmp_path):
    exe = str(tmp_path / "dump_probe")
    out = subprocess.run(
        ["g++", "-std=c++11", "-O1", "-DB200CONV_WITH_STATE_DUMPER", "-DLSP_DSP_UNITS_BUILTIN",
         "-I", os.path.join(HOST, "include"), "-I", os.path.join(ge.ROOT, "include"),
         "-I", REF_INC, "-I", os.path.join(ge.ROOT, "oracle", "shim"),
         os.path.join(HOST, "dump_probe.cpp"), os.path.join(HOST, "Convolver.cpp"),
         "/root/reference/src/main/iface/IStateDumper.cpp",
         "-L", ge.PKG_DIR, "-lb200conv", "-Wl,-rpath," + ge.PKG_DIR, "-o", exe],
        capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="needs the reference headers")
def test_facade_dump_runs_and_writes_the_reference_fields_in_order(tmp_path):
    """dump() is RUN through the reference's IStateDumper: the first 18 writes carry the names of
    Convolver.cpp:317-336 in the reference's order (taken from the verbatim build of the reference
    class), engine extras follow.  Un-initialised object: no GPU needed."""
    from oracle.bindings import CpuConvolver
    ge.load().build()
    exe = _build_dump_probe(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    fields = [ln.split("=") for ln in out.stdout.split()]
    n_ref, ref_names = CpuConvolver("reference").dump_names()
    assert n_ref == 18
    assert [f[0] for f in fields[:18]] == ref_names
    # construct(): every pointer NULL, every scalar 0 (Convolver.cpp:46-69)
    assert all(v in ("null", "0") for _, v in fields[:18])
    assert [f[0] for f in fields[18:]] == ["pEngine", "nDevice", "nPartitions", "nPartOffset", "nFrames"]


@pytest.mark.gpu
def test_host_utest_on_gpu():
    exe = os.path.join(HOST, "utest_convolver")
    if not os.path.exists(exe):
        _make()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASSED" in out.stdout
