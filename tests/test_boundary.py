"""The drop-in boundary: the C-ABI library loads (no GPU needed), exports every symbol that
include/torchrl_b200.h declares, and the product package never imports the oracle."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "torchrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native_lib):
    from torchrl_b200 import _lib
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(native_lib, name), "header declares %s but the library does not export it" % name
    # and the python binding types exactly the declared set
    assert sorted(_lib.SIGNATURES) == declared


def test_abi_version_and_error_string(native_lib):
    assert native_lib.trl_abi_version() == 3
    assert native_lib.trl_last_error() is not None


def test_argument_errors_without_gpu(native_lib):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    rc = native_lib.trl_gae_scan(None, None, None, None, None, None, None, 4, 4, 0.99, 0.95, 1, 1, None)
    assert rc == -1
    assert b"null" in native_lib.trl_last_error()
    rc = native_lib.trl_gae_scan(None, None, None, None, None, None, None, -1, 4, 0.99, 0.95, 1, 1, None)
    assert rc == -1
    rc = native_lib.trl_gae_scan(None, None, None, None, None, None, None, 0, 4, 0.99, 0.95, 1, 1, None)
    assert rc == 0  # empty rollout is a no-op


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torchrl_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                txt = open(os.path.join(dp, fn)).read()
                assert not pat.search(txt), "%s imports the oracle" % os.path.join(dp, fn)
                assert "/root/reference" not in txt.replace("/root/reference/torchrl", "REFDOC"), fn


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from torchrl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setenv("TORCHRL_B200_NO_AUTOBUILD", "1")
    import pytest
    with pytest.raises(_lib.NativeLibraryError):
        _lib.load()


def test_deferred_reduce_scope_is_strict():
    """networks.fused.deferred_reduces(): pending slab-sum jobs must be flushed inside the scope (a silent drop would
    lose gradients); an empty flush is a no-op and the scope restores the previous state.  Host logic only."""
    from torchrl_b200.networks import fused
    assert fused._DEFER is None
    with fused.deferred_reduces():
        assert fused._DEFER == []
        fused.flush_reduces()                      # nothing recorded: no library call
        with fused.deferred_reduces():             # nested scopes keep their own job lists
            assert fused._DEFER == []
        assert fused._DEFER == []
    assert fused._DEFER is None
    with pytest.raises(RuntimeError):
        with fused.deferred_reduces():
            fused._DEFER.append((0, None, None, None, 1, 32, 1, 0))
    assert fused._DEFER is None
    assert not fused._can_defer(object())          # outside a scope nothing is deferred
