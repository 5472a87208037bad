"""CPU: the C-ABI library builds, loads, and exports every symbol include/cdk.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cd_dynamax_b200 import build as b
    b.build()
    from cd_dynamax_b200 import _lib
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cdk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"CDK_DECL\((cdk_\w+)\)", src))
    names |= set(re.findall(r"\b(cdk_\w+)\s*\(", src))
    names -= {"cdk_desc", "cdk_stream_t", "cdk_xla_opaque"}
    return sorted(names)


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 12 + 9
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cdk.h but not exported by libcdk.so"


def test_python_symbol_lists_match_header():
    from cd_dynamax_b200 import _lib
    assert sorted(_lib.ENTRY_POINTS + _lib.OTHER_SYMBOLS) == declared_symbols()


def test_desc_layout_and_defaults(lib):
    from cd_dynamax_b200 import _lib
    d = _lib.new_desc()
    assert d.struct_size == ctypes.sizeof(_lib.CdkDesc)
    assert d.solver == _lib.SOLVERS["dopri5"] and d.dt0 == 0.01 and d.max_steps == 100000  # diffrax_utils.py:50-52
    assert d.dt_final == 1e-10 and d.state_order == 2 and d.num_iter == 1
    assert abs(d.alpha - 3 ** 0.5) < 1e-15 and d.beta == 2.0 and d.kappa == 1.0


def test_invalid_descriptors_are_rejected_on_host(lib):
    """Validation is synchronous and happens before any CUDA call, so it is testable without a GPU."""
    from cd_dynamax_b200 import _lib
    ins = (ctypes.c_void_p * _lib.NUM_IN)()
    outs = (ctypes.c_void_p * _lib.NUM_OUT)()
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m = 4, 10, 3, 1
    d.struct_size = 8
    assert lib.cdk_ekf_filter_f64(ctypes.byref(d), ins, outs, None) == -2  # CDK_E_SIZE
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m, d.solver = 4, 10, 3, 1, 99
    assert lib.cdk_kf_filter_f64(ctypes.byref(d), ins, outs, None) == -3  # CDK_E_ENUM
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m = 4, 10, 3, 1
    assert lib.cdk_kf_filter_f64(ctypes.byref(d), ins, outs, None) == -1  # CDK_E_NULL: inputs missing
    assert b"NULL" in lib.cdk_last_error()
    d.N = 0
    assert lib.cdk_kf_filter_f64(ctypes.byref(d), ins, outs, None) == 0  # empty batch is a no-op
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m, d.drift_id, d.n_theta = 4, 10, 4, 1, 1, 3
    assert lib.cdk_ekf_filter_f64(ctypes.byref(d), ins, outs, None) == -2  # lorenz63 needs n == 3


def test_settings_mapping():
    from cd_dynamax_b200 import _engine as E, solvers
    assert E.parse_settings({}) == {"solver": 6, "dt0": 0.01, "max_steps": 100000}
    assert E.parse_settings({}, sde=True)["solver"] == 1
    assert E.parse_settings({"solver": solvers.RK4(), "dt0": 0.0025, "max_steps": 1e3}) == {
        "solver": 5, "dt0": 0.0025, "max_steps": 1000}
    assert E.parse_settings({"solver": "Euler", "stepsize_controller": solvers.ConstantStepSize()})["solver"] == 0
    with pytest.raises(NotImplementedError):
        E.parse_settings({"solver": "tsit5"})

    class PIDController:
        pass
    with pytest.raises(NotImplementedError):
        E.parse_settings({"stepsize_controller": PIDController()})


def test_scratch_bytes_of_the_pushforward_cache(lib):
    """cdk_scratch_bytes is pure host logic: the only device scratch is the optional CD-KF pushforward cache."""
    from cd_dynamax_b200 import _lib
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m, d.solver = 10, 50, 16, 4, _lib.SOLVERS["rk4"]
    for entry in (b"cdk_kf_filter", b"cdk_kf_smooth", b"cdk_ekf_filter", b"cdk_ukf_filter", b"cdk_enkf_filter"):
        assert lib.cdk_scratch_bytes(ctypes.byref(d), entry) == 0  # flag not set
    d.reserved[2] = _lib.FLAG_KEEP_PUSHFORWARD
    want = 10 * 49 * 2 * 16 * 16 * 8  # [N][K-1][2][n][n] doubles
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_filter") == want
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_smooth") == want
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_ekf_filter") == 0
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_ekf_smooth") == 0
    d.solver = _lib.SOLVERS["dopri5"]  # the reference default runs on the warp kernels too (step polynomial): same cache
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_filter") == want
    d.reserved[2] = _lib.FLAG_KEEP_PUSHFORWARD | _lib.FLAG_DIAG_R  # the Woodbury update runs on the generic kernel: no cache
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_filter") == 0
    d.reserved[2] = _lib.FLAG_KEEP_PUSHFORWARD
    d.solver, d.n = _lib.SOLVERS["rk4"], 17
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_filter") == 0
    d.n, d.smoother_type = 16, 2  # the backward-ODE smoother never reads the cache
    assert lib.cdk_scratch_bytes(ctypes.byref(d), b"cdk_kf_smooth") == 0


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/cdk.h must compile as C (no C++ / torch / CUDA types) and every
    declared entry point must link against libcdk.so from a C translation unit."""
    import subprocess
    from cd_dynamax_b200 import _lib
    src = tmp_path / "use_cdk.c"
    src.write_text(
        '#include "cdk.h"\n'
        "int main(void) {\n"
        "  cdk_desc d; cdk_desc_init(&d);\n"
        "  const void* in[CDK_NUM_IN] = {0}; void* out[CDK_NUM_OUT] = {0};\n"
        "  d.N = 0; d.K = 1; d.n = 3; d.m = 1; d.drift_id = CDK_DRIFT_LORENZ63; d.n_theta = 3;\n"
        "  /* N = 0 is a legal no-op: validates the descriptor on the host and returns without touching the GPU */\n"
        "  int rc = cdk_ekf_filter_f64(&d, in, out, (cdk_stream_t)0);\n"
        "  return (rc == CDK_OK && d.struct_size == (int)sizeof(cdk_desc) && cdk_version() >= 1 &&\n"
        "          cdk_scratch_bytes(&d, \"cdk_kf_filter\") == 0) ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "use_cdk"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lcdk", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)


def test_xla_status_returning_adaptor_reports_failures(lib, tmp_path):
    """cdk_xla_custom_call_status maps a non-zero CDK_E_* code to XlaCustomCallStatusSetFailure (SURVEY 8b "Errors").
    The XLA runtime that owns that symbol is absent here, so a 10-line stand-in is compiled and loaded RTLD_GLOBAL; the
    requests below are rejected by host-side validation before any CUDA call."""
    import subprocess
    from cd_dynamax_b200 import _lib
    src = tmp_path / "xla_stub.c"
    src.write_text(
        "#include <string.h>\n#include <stddef.h>\n"
        "typedef struct { int failed; char message[320]; } XlaCustomCallStatus;\n"
        "void XlaCustomCallStatusSetFailure(XlaCustomCallStatus* s, const char* m, size_t n) {\n"
        "  s->failed = 1; if (n > 319) n = 319; memcpy(s->message, m, n); s->message[n] = 0; }\n")
    so = tmp_path / "libxla_stub.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    ctypes.CDLL(str(so), mode=ctypes.RTLD_GLOBAL)

    class Status(ctypes.Structure):
        _fields_ = [("failed", ctypes.c_int), ("message", ctypes.c_char * 320)]

    class Opaque(ctypes.Structure):
        _fields_ = [("entry_point", ctypes.c_char * 32), ("desc", _lib.CdkDesc)]

    bufs = (ctypes.c_void_p * (_lib.NUM_IN + _lib.NUM_OUT))()
    op = Opaque()
    op.entry_point = b"cdk_ekf_filter_f64"
    op.desc = _lib.new_desc()
    op.desc.N, op.desc.K, op.desc.n, op.desc.m, op.desc.solver = 4, 10, 3, 1, 99  # unknown solver
    blob = ctypes.string_at(ctypes.byref(op), ctypes.sizeof(op))
    st = Status()
    lib.cdk_xla_custom_call_status(None, bufs, blob, len(blob), ctypes.byref(st))
    assert st.failed == 1 and b"unknown solver" in st.message and b"(-3)" in st.message
    assert lib.cdk_xla_last_rc() == -3
    op.entry_point = b"cdk_no_such_entry"
    blob = ctypes.string_at(ctypes.byref(op), ctypes.sizeof(op))
    st = Status()
    lib.cdk_xla_custom_call_status(None, bufs, blob, len(blob), ctypes.byref(st))
    assert st.failed == 1 and b"unknown entry point" in st.message
    # the legacy signature cannot report: the code is still retrievable; a too-short opaque is rejected, not read
    lib.cdk_xla_custom_call(None, bufs, blob[:8], 8)
    assert lib.cdk_xla_last_rc() == -2
    # N = 0 is a valid no-op request
    op.entry_point = b"cdk_ekf_filter_f64"
    op.desc = _lib.new_desc()
    op.desc.N, op.desc.K, op.desc.n, op.desc.m, op.desc.drift_id, op.desc.n_theta = 0, 10, 3, 1, 1, 3
    blob = ctypes.string_at(ctypes.byref(op), ctypes.sizeof(op))
    st = Status()
    lib.cdk_xla_custom_call_status(None, bufs, blob, len(blob), ctypes.byref(st))
    assert st.failed == 0 and lib.cdk_xla_last_rc() == 0


def test_user_drift_variant_library_is_not_interposed_by_the_stock_one(lib):
    """A variant built by build_user_drift exports the same symbols as libcdk.so and is loaded next to it; both are linked
    -Bsymbolic so that each binds its own definitions (host-side validation only: no GPU needed)."""
    from cd_dynamax_b200 import _lib
    from cd_dynamax_b200 import build as b
    example = os.path.join(ROOT, "cd_dynamax_b200", "examples", "vdp_drift.cuh")
    var = _lib.lib(b.build_user_drift(open(example).read()))
    assert lib.cdk_has_user_drift() == 0 and var.cdk_has_user_drift() == 1
    ins = (ctypes.c_void_p * _lib.NUM_IN)()
    outs = (ctypes.c_void_p * _lib.NUM_OUT)()
    d = _lib.new_desc()
    d.N, d.K, d.n, d.m, d.drift_id, d.n_theta = 0, 3, 2, 1, _lib.DRIFT_USER, 2
    assert lib.cdk_ekf_filter_f64(ctypes.byref(d), ins, outs, None) == -4  # stock: CDK_E_UNSUPPORTED
    assert var.cdk_ekf_filter_f64(ctypes.byref(d), ins, outs, None) == 0   # variant: valid (N = 0 is a no-op)
