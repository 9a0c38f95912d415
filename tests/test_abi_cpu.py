"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol the header declares,
and the Python binding's struct layouts agree with the C compiler's."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from cswin_unet_b200 import _lib


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def header_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cswin_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ("cswin_lepe_attention_fwd", "cswin_lepe_attention_bwd", "cswin_linear_fwd", "cswin_layernorm_fwd",
              "cswin_im2col_tokens", "cswin_im2col_nchw", "cswin_carafe_reassemble_fwd", "cswin_last_error",
              "cswin_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    for s in header_symbols():
        assert hasattr(built, s), f"libcswin_b200.so does not export {s}"
    assert set(header_symbols()) == set(_lib.SIGNATURES), "binding table and header disagree"
    assert built.cswin_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_c_compiler(built):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "cswin_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cswin_lepe_branch_t), offsetof(cswin_lepe_branch_t, lse),
         offsetof(cswin_lepe_branch_t, W_sp), sizeof(cswin_lepe_branch_grad_t), offsetof(cswin_lepe_branch_grad_t, dconv_b),
         sizeof(cswin_linear_args_t));
  printf("%zu %zu %zu %zu %zu\n", offsetof(cswin_linear_args_t, ln_eps), offsetof(cswin_linear_args_t, rows_per_sample),
         offsetof(cswin_linear_args_t, stats_out), offsetof(cswin_linear_args_t, aux_out), offsetof(cswin_linear_args_t, ld_aux));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cswin_qkv_attn_args_t), offsetof(cswin_qkv_attn_args_t, ln_eps), offsetof(cswin_qkv_attn_args_t, out),
         offsetof(cswin_qkv_attn_args_t, n_branches), offsetof(cswin_qkv_attn_args_t, br), offsetof(cswin_qkv_attn_args_t, scale));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.dirname(_lib.HEADER_PATH), c, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    got = [int(v) for v in out]
    B, G, L, Q = _lib.LepeBranch, _lib.LepeBranchGrad, _lib.LinearArgs, _lib.QkvAttnArgs
    want = [ctypes.sizeof(B), B.lse.offset, B.W_sp.offset, ctypes.sizeof(G), G.dconv_b.offset, ctypes.sizeof(L),
            L.ln_eps.offset, L.rows_per_sample.offset, L.stats_out.offset, L.aux_out.offset, L.ld_aux.offset,
            ctypes.sizeof(Q), Q.ln_eps.offset, Q.out.offset, Q.n_branches.offset, Q.br.offset, Q.scale.offset]
    assert got == want


def test_errors_are_reported_not_raised_across_the_abi(built):
    # invalid arguments must come back as an error code + message without touching the GPU
    rc = built.cswin_lepe_attention_fwd(None, 3, 1, 56, ctypes.c_float(1.0), 0, None)
    assert rc == 1 and b"n_branches" in built.cswin_last_error()
    rc = built.cswin_linear_fwd(None, 7, None)
    assert rc == 1 and b"dtype" in built.cswin_last_error()
    with pytest.raises(_lib.CswinError):
        _lib.check(rc, "cswin_linear_fwd")
