"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol
include/car_b200.h declares, the ctypes structs match the C layout, and argument errors are
reported without touching a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from cross_attention_renderer_b200 import _lib

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(REPO, "include", "car_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(car_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    # the micro-benchmark / building-block kernels live in their own library and header
    tsrc = open(os.path.join(REPO, "include", "car_b200_test.h")).read()
    tdecl = set(re.findall(r"\b(car_[a-z0-9_]+)\s*\(", tsrc))
    assert tdecl == set(_lib.TEST_SYMBOLS), (tdecl ^ set(_lib.TEST_SYMBOLS))
    tlib = _lib.load_test()
    for name in tdecl:
        assert hasattr(tlib, name), name
        assert not hasattr(lib, name), f"{name} must not be exported by the product library"


def test_version(lib):
    assert lib.car_version() == _lib.ABI_VERSION
    m = re.search(r"#define CAR_ABI_VERSION (\d+)", open(HEADER).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_struct_layout_matches_c(tmp_path):
    """Compile a C probe against the header and compare sizeof/offsetof with ctypes."""
    probe = tmp_path / "probe.c"
    probe.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "car_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(car_mat), sizeof(car_weights), sizeof(car_cameras),
         sizeof(car_debug), sizeof(car_render_args), sizeof(car_weight_grads), sizeof(car_backward_args),
         offsetof(car_render_args, train), offsetof(car_render_args, chunk_rays));
  printf("%zu %zu %zu %zu\\n", offsetof(car_backward_args, fwd), offsetof(car_backward_args, grads),
         offsetof(car_backward_args, d_feat), offsetof(car_backward_args, stream));
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", offsetof(car_render_args, feat), offsetof(car_render_args, weights),
         offsetof(car_render_args, cams), offsetof(car_render_args, uv),
         offsetof(car_render_args, workspace_bytes), offsetof(car_render_args, stream),
         offsetof(car_render_args, use_fused));
  return 0;
}''')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(probe), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    sizes = list(map(int, lines[0].split()))
    boffs = list(map(int, lines[1].split()))
    offs = list(map(int, lines[2].split()))
    A = _lib.car_render_args
    B = _lib.car_backward_args
    assert sizes == [C.sizeof(_lib.car_mat), C.sizeof(_lib.car_weights), C.sizeof(_lib.car_cameras),
                     C.sizeof(_lib.car_debug), C.sizeof(A), C.sizeof(_lib.car_weight_grads), C.sizeof(B),
                     A.train.offset, A.chunk_rays.offset]
    assert boffs == [B.fwd.offset, B.grads.offset, B.d_feat.offset, B.stream.offset]
    assert offs == [A.feat.offset, A.weights.offset, A.cams.offset, A.uv.offset,
                    A.workspace_bytes.offset, A.stream.offset, A.use_fused.offset]


def test_sizes_and_argument_errors(lib):
    assert lib.car_features_bytes(2, 64, 64, 0, 0) == 2 * 16 * 16 * 256 * 4
    assert lib.car_features_bytes(2, 64, 64, 2, 1) == 2 * 64 * 64 * 64 * 2
    c = lib.car_default_chunk_rays(0, 64, 0)
    assert c >= 1
    assert lib.car_workspace_bytes(0, 64, c, 0) > lib.car_workspace_bytes(0, 64, 1, 0) > 0
    # the fused path never materialises the 576-wide activations
    assert lib.car_workspace_bytes(1, 64, 1024, 3) < lib.car_workspace_bytes(1, 64, 1024, 0) / 4
    assert lib.car_default_chunk_rays(1, 64, 3) > lib.car_default_chunk_rays(1, 64, 0)
    a = _lib.car_render_args()
    a.abi_version = 1
    assert lib.car_render_forward(C.byref(a)) == -2
    assert b"ABI" in lib.car_last_error()
    a.abi_version = _lib.ABI_VERSION
    assert lib.car_render_forward(C.byref(a)) == -3          # sizes are all zero
    assert lib.car_render_forward(None) == -1
    # training / backward sizing and argument checks
    assert lib.car_train_workspace_bytes(0, 64, 192) > lib.car_workspace_bytes(0, 64, 192, 0)
    assert lib.car_backward_workspace_bytes(0, 64, 192) > lib.car_backward_workspace_bytes(0, 64, 1) > 0
    assert lib.car_backward_workspace_bytes(1, 64, 192) > lib.car_backward_workspace_bytes(0, 64, 192)
    assert lib.car_train_workspace_bytes(1, 64, 192) > lib.car_train_workspace_bytes(0, 64, 192)
    bw = _lib.car_backward_args()
    assert lib.car_render_backward(None) == -1
    assert lib.car_render_backward(C.byref(bw)) == -1        # no forward arguments
    bw.fwd = C.pointer(a)
    bw.abi_version = _lib.ABI_VERSION
    assert lib.car_render_backward(C.byref(bw)) == -12       # forward was not a train=1 call
    assert b"train" in lib.car_last_error()


def test_no_cpu_fallback():
    import torch
    from cross_attention_renderer_b200 import synthetic
    from cross_attention_renderer_b200.models import CrossAttentionRenderer
    m = CrossAttentionRenderer(n_view=2, npoints=8)
    inp = synthetic.make_inputs(1, 32, 8)
    z = synthetic.make_features(1, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(inp, z=z)
