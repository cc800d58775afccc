"""The C-ABI boundary without a GPU: libunetk.so loads, exports every symbol include/unetk.h declares,
and the ctypes signatures in jcfszxc_unet_b200/_lib.py agree with the header (count and C type class).
No compute entry point is called here."""
import ctypes as C
import re

import pytest

from jcfszxc_unet_b200 import _lib


def _header_decls():
    text = re.sub(r"/\*.*?\*/", "", _lib.HEADER.read_text(), flags=re.S)
    text = re.sub(r"#.*", "", text)
    decls = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(unetk_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("void", "") else [p.strip() for p in params.split(",")]
        decls[name] = (ret, plist)
    return decls


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return C.c_void_p
    base = decl.split()[:-1] if len(decl.split()) > 1 else decl.split()
    base = " ".join(b for b in base if b != "const")
    return {"int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "float": C.c_float, "double": C.c_double}[base]


def test_library_builds_and_loads():
    lib = _lib.load()
    assert lib.unetk_abi_version() == 3
    assert lib.unetk_last_error() is not None


def test_every_header_symbol_is_exported_and_bound():
    decls = _header_decls()
    assert set(decls) == set(_lib.header_symbols())
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in decls:
        assert hasattr(lib, name), f"{name} not exported by libunetk.so"


@pytest.mark.parametrize("name", sorted(_lib.SIGNATURES))
def test_signature_matches_header(name):
    ret, params = _header_decls()[name]
    res, args = _lib.SIGNATURES[name]
    assert len(params) == len(args), f"{name}: header has {len(params)} params, binding {len(args)}"
    for i, (p, a) in enumerate(zip(params, args)):
        assert _ctype_of(p) is a, f"{name} arg {i} ({p}): binding says {a}"
    if ret == "int":
        assert res is C.c_int
    elif ret == "size_t":
        assert res is C.c_size_t
    elif ret == "const char*" or ret.replace(" ", "") == "constchar*":
        assert res is C.c_char_p


def test_size_queries_are_pure_host_functions():
    lib = _lib.load()
    assert lib.unetk_conv_wgrad_workspace(2, 64, 64, 64, 64, 9) > 0
    assert lib.unetk_chan_partial_floats(4096, 64) > 0
    assert lib.unetk_head_partial_floats(4096, 64) > 0
    assert lib.unetk_sqnorm_partial_floats(1 << 20) > 0
    assert lib.unetk_stem_wgrad_workspace(2, 64, 64, 3) > 0
