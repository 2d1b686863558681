"""CPU: the C-ABI library loads and exports every symbol include/b200conformer.h declares (no compute calls)."""
import ctypes


def test_library_loads_and_exports_all_declared_symbols():
    from nemo_b200 import _lib

    syms = _lib.declared_symbols()
    assert len(syms) >= 25
    for name, _res, _args in syms:
        assert hasattr(_lib.lib, name), f"{name} declared in include/b200conformer.h but not exported"
    assert _lib.lib.b2c_version() >= 100
    assert isinstance(_lib.last_error(), str)
    assert _lib.lib.b2c_launch_count() == 0


def test_no_undeclared_exports():
    import subprocess

    from nemo_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T b2c_" in l}
    declared = {n for n, _, _ in _lib.declared_symbols()}
    assert exported == declared, (exported - declared, declared - exported)


def test_argument_validation_does_not_need_a_gpu():
    from nemo_b200 import _lib

    e = _lib.Epilogue()
    rc = _lib.lib.b2c_gemm(None, None, 0, 0, 0, 0, 0, 0, 0, 0, ctypes.byref(e), None)
    assert rc == -1 and "bad shape" in _lib.last_error()


def test_more_argument_validation_without_a_gpu():
    """error behaviour of the entry points added this round: a negative return code + a message, before any CUDA call"""
    from nemo_b200 import _lib

    # dropout epilogues need N % 16 == 0 (one hash chunk per 16 output columns)
    e = _lib.Epilogue()
    e.out0 = ctypes.c_void_p(16)
    e.drop_keep = 59000
    rc = _lib.lib.b2c_gemm(ctypes.c_void_p(16), ctypes.c_void_p(16), 128, 24, 64, 64, 64, 0, 0, 0, ctypes.byref(e), None)
    assert rc < 0 and "N % 16" in _lib.last_error()
    rc = _lib.lib.b2c_spec_augment(None, None, None, None, None, None, 0, 80, 100, 1, 1, ctypes.c_float(0.0), None)
    assert rc < 0 and "spec_augment" in _lib.last_error()
    rc = _lib.lib.b2c_subsample_im2col(None, None, None, 0, 80, 100, None)
    assert rc < 0 and "im2col" in _lib.last_error()
    assert _lib.lib.b2c_launch_count() == 0  # nothing was launched
