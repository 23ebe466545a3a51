"""Writes tests/golden/qr_lstsq_example.json: the reference's only known-answer vector on the QR path.
Source of the numbers: /root/reference/faer/src/linalg/qr/mod.rs:123-150 (`test_example`: a 10x2 least-squares
problem whose expected solution was produced with numpy; the reference checks |x - expected| <= 1e-6).
Run here (needs /root/reference); the JSON is committed because /root/reference does not exist on the GPU box."""
import json, os, re
src = open("/root/reference/faer/src/linalg/qr/mod.rs").read()
def grab(name):
    m = re.search(r"let %s = mat!\[(.*?)\];" % name, src, re.S)
    rows = re.findall(r"\[([^\[\]]+)\]", m.group(1))
    return [[float(x.replace("_f64", "")) for x in r.split(",") if x.strip()] for r in rows]
out = {"source": "faer/src/linalg/qr/mod.rs:123-150", "tolerance": 1e-6,
       "a": grab("a"), "b": grab("b"), "expected_solution": grab("expected_solution")}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "qr_lstsq_example.json"), "w"), indent=1)
print({k: (len(v), len(v[0])) for k, v in out.items() if isinstance(v, list)})
