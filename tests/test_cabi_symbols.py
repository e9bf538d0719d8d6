"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        src = open(os.path.join(inc, fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms += re.findall(r"\b((?:rs|kb)_[a-z_0-9]+)\s*\(", src)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    from ranslice_b200 import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert "rs_create" in syms and "rs_step" in syms and "rs_step_device" in syms
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s


def test_create_rejects_bad_arguments_without_gpu():
    """Argument validation happens before any CUDA call (error behaviour of the boundary)."""
    from ranslice_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.rs_create(None, None, ctypes.byref(h)) == -1
    cfg = _lib.RsConfig(abi_version=999)
    tb = _lib.RsTables()
    assert L.rs_create(ctypes.byref(cfg), ctypes.byref(tb), ctypes.byref(h)) == -1
    assert b"abi_version" in L.rs_last_error()
    assert L.rs_reset(None, None) == -1 and L.rs_step(None, None, None, None, None, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from ranslice_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libranslice_b200.so")
    try:
        _lib.lib()
    except _lib.NativeLibraryMissing as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected NativeLibraryMissing")


def test_controller_and_wrapper_entry_points_validate_arguments_without_gpu():
    """kb_* / rs_wrap_* / rs_step_async reject bad arguments before any CUDA call (error behaviour of the boundary)."""
    from ranslice_b200 import _lib
    L = _lib.lib()
    vp = ctypes.c_void_p
    assert L.kb_create(None, None, None, None) == -1
    assert L.kb_control_init(None, None, None, ctypes.c_double(0.05), ctypes.c_double(0.97), ctypes.c_double(0.99)) == -1
    assert L.kb_control_update_device(None, None, None, None, None, None) == -1
    assert L.kb_control_select_device(None, None, None, None, None) == -1
    assert L.kb_set_exact(None, 1) == -1
    assert L.rs_wrap_action_device(None, 0, 4, 5, 200, None, None) == -1
    buf = (ctypes.c_int32 * 64)()
    p = ctypes.cast(buf, vp)
    assert L.rs_wrap_action_device(p, 0, 4, 9, 200, p, None) == -1            # more than 8 slices
    assert b"n_slices" in L.rs_last_error()
    assert L.rs_wrap_obs_device(p, p, ctypes.c_int64(0), None) == -1
    assert L.rs_wrap_record_device(p, p, p, 4, 5, ctypes.c_int64(-1), None, None, None, None) == -1
    t = ctypes.c_int32()
    assert L.rs_step_async(None, None, None, None, None, None, None, ctypes.byref(t)) == -1
    assert L.rs_wait(None, 0) == -1
