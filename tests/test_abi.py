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
