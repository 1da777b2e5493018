"""kjx_xla_ffi.cc (the XLA-FFI custom-call handlers over the C ABI) must stay a compilable translation unit even though
jaxlib's headers are not in this image: compile it against tests/xla_stub/, a minimal stand-in for xla/ffi/api/ffi.h whose
handler macro also checks every handler's signature against its binding."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "kalman_jax_b200", "csrc", "kjx_xla_ffi.cc")
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="needs g++ and the CUDA headers")
def test_xla_ffi_handlers_compile_and_match_their_bindings(tmp_path):
    obj = tmp_path / "kjx_xla_ffi.o"
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "tests", "xla_stub"), "-I", CUDA_INC,
           SRC, "-o", str(obj)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    syms = subprocess.run(["nm", "-g", "--defined-only", str(obj)], capture_output=True, text=True).stdout
    assert "KjxRunF64" in syms and "KjxSiteUpdateF64" in syms        # not compiled to nothing


def test_a_handler_that_disagrees_with_its_binding_is_rejected(tmp_path):
    """The stub's macro is a real check: drop one argument from a binding and compilation must fail."""
    bad = tmp_path / "bad.cc"
    src = open(SRC).read().replace('.Attr<ffi::Span<const double>>("h")\n                                  .Arg<ffi::Buffer<ffi::F64>>()   // dt',
                                   '.Arg<ffi::Buffer<ffi::F64>>()   // dt')
    assert src != open(SRC).read()
    bad.write_text(src.replace('#include "../../include/kjx.h"', '#include "%s"' % os.path.join(ROOT, "include", "kjx.h")))
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "xla_stub"), "-I", CUDA_INC,
                          str(bad)], capture_output=True, text=True)
    assert out.returncode != 0 and "does not take the arguments its binding declares" in out.stderr
