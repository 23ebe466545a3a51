"""Generator of tests/golden/faer_ffi_prototypes.json: the C prototypes (return type + parameter types, names dropped) that the
reference's shipped header `faer-ffi/faer.h` declares for the `libfaer_v0_23_*` symbols this repo exports, plus the full list
of symbol names the reference declares for f32 / f64 / c32 / c64 (the coverage denominator of DESIGN.md section 1).
Run in the build container (the reference is not present on the GPU box): python tests/golden/make_ffi_prototypes.py
tests/test_cabi_cpu.py::test_prototypes_match_the_reference_header compares include/faer_b200.h against the committed file."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/faer-ffi/faer.h"


def prototypes(text: str) -> dict:
    """name -> [return type, [parameter types]] for every `libfaer_v0_23_*` declaration in a C header."""
    out = {}
    norm = lambda x: re.sub(r"\s+", " ", x.replace("struct ", "").replace("enum ", "")).strip().replace(" *", "*")
    for m in re.finditer(r"([A-Za-z_0-9 \*]+?)\s*\b(libfaer_v0_23_[A-Za-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, re.S):
        ret, name, args = m.groups()
        types = []
        for a in args.split(","):
            a = norm(a)
            if a in ("void", ""):
                continue
            if not a.endswith("*"):
                a = re.sub(r"\s*\b[A-Za-z_][A-Za-z0-9_]*$", "", a)  # drop the parameter name
            types.append(a)
        out[name] = [norm(ret), types]
    return out


def main() -> int:
    ref = prototypes(open(REF).read())
    ours = prototypes(open(os.path.join(ROOT, "include", "faer_b200.h")).read())
    in_scope = sorted(n for n in ref if not re.search(r"(fx128|cx128)$", n))
    doc = {
        "source": "faer-ffi/faer.h (reference), parsed by tests/golden/make_ffi_prototypes.py",
        "prototypes": {n: ref[n] for n in sorted(ours) if n in ref},
        "reference_symbols_f32_f64_c32_c64": in_scope,
    }
    path = os.path.join(ROOT, "tests", "golden", "faer_ffi_prototypes.json")
    json.dump(doc, open(path, "w"), indent=0, sort_keys=True)
    missing = [n for n in ours if n not in ref]
    print(f"{len(doc['prototypes'])} prototypes written, {len(in_scope)} reference symbols in the four dtypes; "
          f"declared here but not in the reference: {missing}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
