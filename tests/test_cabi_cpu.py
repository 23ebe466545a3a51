"""CPU-side tests of the drop-in boundary (no compute calls: there is no CPU fallback by design).

* libfaer_b200.so loads and exports every symbol include/faer_b200.h declares;
* POD layouts match faer-ffi's (sizes/offsets as in /root/reference/faer-ffi/faer.h: 5-field views, by-value params,
  tagged-union statuses);
* default params mirror the reference (LLT {64,128}: ldlt/factor.rs:705-714; LU {16,64,16384}: lu/.../factor.rs:212-222);
* alloc/dealloc and global-par entry points behave like the reference's (faer-ffi/src/lib.rs:2521-2569);
* the product does NOT import or link the oracle.
"""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "faer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b((?:libfaer_v0_23|faer_b200)_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(fb):
    lib = fb.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/faer_b200.h but not exported by libfaer_b200.so"
    out = subprocess.check_output(["nm", "-D", "--defined-only", fb.capi.LIB_PATH], text=True)
    exported = set(re.findall(r"\b((?:libfaer_v0_23|faer_b200)_\w+)", out))
    assert set(syms) <= exported


def test_pod_layouts_match_faer_ffi(fb):
    capi = fb.capi
    assert C.sizeof(capi.MatRef) == 40 and C.sizeof(capi.MatMut) == 40
    assert [f[0] for f in capi.MatRef._fields_] == ["ptr", "nrows", "ncols", "row_stride", "col_stride"]
    assert C.sizeof(capi.Par) == 16 and capi.Par.nthreads.offset == 8
    assert C.sizeof(capi.Layout) == 16 and C.sizeof(capi.MemAlloc) == 16 and C.sizeof(capi.SliceMut) == 16
    assert C.sizeof(capi.LltParams) == 16 and C.sizeof(capi.PartialPivLuParams) == 24
    # tagged unions: 4-byte tag padded to 8, then one size_t body (faer.h:383-469)
    for st in (capi.LltStatus, capi.PartialPivLuStatus, capi.QrStatus):
        assert C.sizeof(st) == 16 and st.body.offset == 8


def test_default_params_and_helpers(fb):
    lib = fb.load()
    p = lib.libfaer_v0_23_LltParams_f64()
    assert (p.recursion_threshold, p.block_size) == (64, 128)
    q = lib.libfaer_v0_23_PartialPivLuParams_f64()
    assert (q.recursion_threshold, q.block_size, q.par_threshold) == (16, 64, 128 * 128)
    par = fb.capi.par_default()
    lay = lib.libfaer_v0_23_llt_factor_in_place_scratch_f64(1000, par, p)
    assert lay.len_bytes == 1000 * 8  # temp_mat_scratch::<f64>(dim, 1), llt/factor.rs:58-66
    lay = lib.libfaer_v0_23_partial_piv_lu_factor_in_place_scratch_u64_f64(300, 200, par, q)
    assert lay.len_bytes == 200 * 8
    lay = lib.libfaer_v0_23_partial_piv_lu_factor_in_place_scratch_u32_f64(300, 200, par, q)
    assert lay.len_bytes == 200 * 4
    # global par round trip
    lib.libfaer_v0_23_set_global_par(fb.capi.Par(fb.capi.PAR_SEQ, 0))
    assert lib.libfaer_v0_23_get_global_par().tag == fb.capi.PAR_SEQ
    lib.libfaer_v0_23_set_global_par(fb.capi.Par(fb.capi.PAR_RAYON, 8))
    g = lib.libfaer_v0_23_get_global_par()
    assert g.tag == fb.capi.PAR_RAYON and g.nthreads == 8
    # alloc honours alignment
    ptr = lib.libfaer_v0_23_alloc(1000, 128)
    assert ptr and ptr % 128 == 0
    lib.libfaer_v0_23_dealloc(ptr, 1000, 128)
    assert b"sm_100a" in lib.faer_b200_version()


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "faer-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle: the product path must not depend on it"
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libfaer_b200.so")], text=True)
    assert "oracle" not in out


def test_qr_solve_scratch_queries(fb):
    """qr/no_pivoting/solve.rs:3-37: all three solves need the block-Householder sequence scratch,
    temp_mat_scratch(block_size, rhs_ncols)."""
    lib = fb.load()
    par = fb.capi.par_default()
    for suf, sz in (("f64", 8), ("f32", 4)):
        lay = getattr(lib, f"libfaer_v0_23_qr_solve_lstsq_in_place_scratch_{suf}")(1000, 300, 32, 7, par)
        assert lay.len_bytes == 32 * 7 * sz
        for name in ("qr_solve_in_place", "qr_solve_transpose_in_place"):
            lay = getattr(lib, f"libfaer_v0_23_{name}_scratch_{suf}")(300, 16, 5, par)
            assert lay.len_bytes == 16 * 5 * sz


def test_prototypes_match_the_reference_header():
    """Every `libfaer_v0_23_*` prototype of include/faer_b200.h has the return type and parameter types that the reference's
    shipped header declares for the same symbol (tests/golden/faer_ffi_prototypes.json, extracted from faer-ffi/faer.h by
    tests/golden/make_ffi_prototypes.py with the same parser), and nothing is declared that the reference does not have."""
    import importlib.util
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "faer_ffi_prototypes.json")))
    spec = importlib.util.spec_from_file_location("make_ffi_prototypes", os.path.join(ROOT, "tests", "golden", "make_ffi_prototypes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ours = mod.prototypes(open(os.path.join(ROOT, "include", "faer_b200.h")).read())
    assert len(ours) >= 250
    for name, proto in ours.items():
        assert name in gold["prototypes"], f"{name} is not a faer-ffi symbol"
        assert gold["prototypes"][name] == proto, (name, gold["prototypes"][name], proto)
    assert set(ours) <= set(gold["reference_symbols_f32_f64_c32_c64"])


def test_reconstruct_inverse_scratch_queries_every_dtype(fb):
    """`*_reconstruct_scratch` / `*_inverse_scratch` for f64 / f32 / c64 / c32: the reference's formulas with the element size
    of T (llt/reconstruct.rs:3-6 EMPTY, llt/inverse.rs:3-8 temp_mat(dim, dim), lu/partial_pivoting/reconstruct.rs
    temp_mat(nrows, ncols), inverse.rs:3-10 temp_mat(dim, dim), qr/no_pivoting/reconstruct.rs:3-12 and inverse.rs:3-10: the
    block-Householder sequence scratch temp_mat(block_size, ncols))."""
    lib = fb.load()
    par = fb.capi.par_default()
    for suf, sz in (("f64", 8), ("f32", 4), ("c64", 16), ("c32", 8)):
        assert getattr(lib, f"libfaer_v0_23_llt_reconstruct_scratch_{suf}")(100, par).len_bytes == 0
        assert getattr(lib, f"libfaer_v0_23_llt_inverse_scratch_{suf}")(100, par).len_bytes == 100 * 100 * sz
        for it in ("u32", "u64"):
            assert getattr(lib, f"libfaer_v0_23_partial_piv_lu_reconstruct_scratch_{it}_{suf}")(30, 70, par).len_bytes == 30 * 70 * sz
            assert getattr(lib, f"libfaer_v0_23_partial_piv_lu_inverse_scratch_{it}_{suf}")(50, par).len_bytes == 50 * 50 * sz
        assert getattr(lib, f"libfaer_v0_23_qr_reconstruct_scratch_{suf}")(300, 120, 32, par).len_bytes == 32 * 120 * sz
        assert getattr(lib, f"libfaer_v0_23_qr_inverse_scratch_{suf}")(90, 16, par).len_bytes == 16 * 90 * sz


def test_state_touching_entry_points_are_serialised(fb):
    """The entry points that touch per-process state take the library's entry lock (runtime.cuh: FB_ENTRY); hammer the
    ones that need no GPU from several threads (ctypes drops the GIL during the calls)."""
    import threading
    lib = fb.load()
    errs = []

    def work():
        try:
            for _ in range(2000):
                lib.faer_b200_set_stream(None)
                lib.faer_b200_launch_count()
                lib.faer_b200_release_workspace()
                p = lib.libfaer_v0_23_alloc(256, 64)
                assert p and p % 64 == 0
                lib.libfaer_v0_23_dealloc(p, 256, 64)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
