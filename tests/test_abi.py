"""The C-ABI boundary: libb200rl.so loads and exports every symbol include/b200rl.h declares, the
ctypes prototypes cover them all, and argument errors come back as codes + messages (no compute:
nothing here needs a GPU)."""
import ctypes as C
import os
import re

import pytest

from rl_algo_impls_b200 import _lib

HEADER = os.path.join(_lib.REPO_ROOT, "include", "b200rl.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_prototypes_agree():
    syms = declared_symbols()
    assert len(syms) >= 19
    assert set(syms) == set(_lib.PROTOTYPES), set(syms) ^ set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `make` / __graft_entry__.build() first"
    handle = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} is declared in b200rl.h but not exported"


def test_version_and_error_channel():
    L = _lib.lib()
    assert L.b200rl_version() == 102
    # a null pointer is an argument error, reported through the code + last_error, never a crash
    rc = L.b200rl_gae_scan_f32(None, None, None, None, None, None, None, 1, None, None, 4, 4, 1, None)
    assert rc == -1
    assert b"null pointer" in L.b200rl_last_error()
    with pytest.raises(_lib.B200RLError, match="null pointer"):
        _lib.check(rc, "b200rl_gae_scan_f32")


def test_workspace_queries_are_host_only():
    L = _lib.lib()
    assert L.b200rl_ppo_workspace_bytes(3072, 1) >= 3072 * 6 * 8
    assert L.b200rl_ppo_workspace_bytes(128, 13) >= 128 * 30 * 8
    assert L.b200rl_adv_moments_workspace_bytes(3072, 13) > 0


def test_structs_match_the_header_layout():
    # field order of the structs that cross the ABI (a reordering would silently corrupt calls)
    text = open(HEADER).read()
    for struct, cls in (("b200rl_ppo_args", _lib.PpoArgs), ("b200rl_gridnet_desc", _lib.GridnetDesc),
                        ("b200rl_store_pack", _lib.StorePack)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]
        assert names == [f[0] for f in cls._fields_], (struct, names)
