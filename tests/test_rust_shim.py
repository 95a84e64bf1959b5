"""The Rust shim crate (madsim-b200/) cannot be compiled here — no cargo / rustc in the image — so its FFI layer is
checked as TEXT: every `extern "C"` item of src/ffi.rs against the prototype in include/msim_b200.h (name, argument
count, each argument's type, return type) and every #[repr(C)] struct against the C struct (field names, order, types).
Reference seam the crate stands in for: madsim/src/sim/runtime/builder.rs:121-162."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "msim_b200.h")).read()
FFI = open(os.path.join(ROOT, "madsim-b200", "src", "ffi.rs")).read()
LIB = open(os.path.join(ROOT, "madsim-b200", "src", "lib.rs")).read()


def strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


# C type (normalised) -> Rust type
C2R = {
    "int": "c_int", "void": "()", "double": "f64", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "size_t": "usize",
    "const char*": "*const c_char", "void*": "*mut c_void", "void**": "*mut *mut c_void", "size_t*": "*mut usize",
    "uint64_t*": "*mut u64", "uint32_t*": "*mut u32", "uint8_t*": "*mut u8",
}


def c_type_to_rust(t):
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*", "*", t)
    if t in C2R:
        return C2R[t]
    m = re.fullmatch(r"(const )?(msim_[a-z_]+)(\*+)", t)
    if m:
        const, name, stars = m.groups()
        out = name
        for i, _ in enumerate(stars):
            out = ("*const " if (const and i == 0) else "*mut ") + out
        return out
    m = re.fullmatch(r"(msim_[a-z_]+)", t)
    if m:
        return t
    raise AssertionError(f"no Rust spelling known for C type {t!r}")


def header_functions():
    src = strip_c_comments(HDR)
    out = {}
    for m in re.finditer(r"^\s*(int|void|const char\*)\s+(msim_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", src, flags=re.M | re.S):
        ret, name, args = m.groups()
        args = args.strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\s+", " ", a.strip())
                mm = re.fullmatch(r"(.*?[\s\*])([A-Za-z_][A-Za-z_0-9]*)", a)
                assert mm, f"cannot split parameter {a!r} of {name}"
                params.append((mm.group(2), c_type_to_rust(mm.group(1))))
        out[name] = (c_type_to_rust(ret), params)
    return out


def rust_functions():
    block = re.search(r'extern "C" \{(.*?)\n\}', FFI, flags=re.S).group(1)
    out = {}
    for m in re.finditer(r"pub fn (msim_[a-z_0-9]+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, args, ret = m.groups()
        params = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            pn, pt = a.split(":", 1)
            params.append((pn.strip(), re.sub(r"\s+", " ", pt.strip())))
        out[name] = ((ret or "()").strip(), params)
    return out


def header_structs():
    src = strip_c_comments(HDR)
    out = {}
    for m in re.finditer(r"typedef struct (msim_[a-z_]+) \{(.*?)\}\s*\1\s*;", src, flags=re.S):
        name, body = m.groups()
        fields = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            decl = re.sub(r"\s+", " ", decl)
            ty, rest = decl.split(" ", 1)
            for f in [x.strip() for x in rest.split(",")]:
                mm = re.fullmatch(r"([A-Za-z_0-9]+)(?:\[(\d+)\])?", f)
                assert mm, f"cannot parse field {f!r} of {name}"
                rt = C2R[ty]
                fields.append((mm.group(1), f"[{rt}; {mm.group(2)}]" if mm.group(2) else rt))
        out[name] = fields
    return out


def rust_structs():
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (msim_[a-z_]+) \{(.*?)\n\}", FFI, flags=re.S):
        name, body = m.groups()
        fields = []
        for f in re.finditer(r"pub ([a-z_0-9]+):\s*([^,\n]+),", body):
            fields.append((f.group(1), f.group(2).strip()))
        out[name] = fields
    return out


def test_every_export_is_bound_with_the_same_signature():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 20
    assert set(h) == set(r), f"header-only: {sorted(set(h) - set(r))}; ffi.rs-only: {sorted(set(r) - set(h))}"
    for name, (ret, params) in h.items():
        rret, rparams = r[name]
        assert rret == ret, f"{name}: return {rret} != {ret}"
        assert [t for _, t in rparams] == [t for _, t in params], f"{name}: {rparams} != {params}"
        assert [n for n, _ in rparams] == [n for n, _ in params], f"{name}: parameter names differ"


def test_repr_c_structs_match_the_header_field_for_field():
    h, r = header_structs(), rust_structs()
    assert {"msim_config", "msim_workload_desc", "msim_seed_result", "msim_trace_rec", "msim_stats"} <= set(h)
    for name, fields in h.items():
        assert name in r, f"ffi.rs has no #[repr(C)] struct {name}"
        assert r[name] == fields, f"{name}: {r[name]} != {fields}"


def test_constants_agree():
    for name, val in re.findall(r"#define (MSIM_(?:OK|E[A-Z]+|ST_[A-Z_]+|WL_[A-Z]+)) \(?(-?\d+)\)?", HDR):
        m = re.search(rf"pub const {name}: (?:c_int|u32) = (-?\d+);", FFI)
        assert m and int(m.group(1)) == int(val), f"{name} missing or different in ffi.rs"


def test_builder_mirrors_the_reference_environment_contract():
    # builder.rs:64-118: the six variables, `count = count.max(2)` under the determinism check, the repro note
    for var in ("MADSIM_TEST_SEED", "MADSIM_TEST_NUM", "MADSIM_TEST_JOBS", "MADSIM_TEST_CONFIG", "MADSIM_TEST_TIME_LIMIT",
                "MADSIM_TEST_CHECK_DETERMINISM"):
        assert var in LIB
    assert "count = count.max(2)" in LIB and "as_nanos() as _" in LIB
    assert "note: run with `MADSIM_TEST_SEED={seed}` environment variable to reproduce this error" in LIB
    # lib.rs calls only functions ffi.rs declares
    used = set(re.findall(r"ffi::(msim_[a-z_0-9]+)\(", LIB))
    assert used <= set(rust_functions())
