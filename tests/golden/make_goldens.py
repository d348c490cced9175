#!/usr/bin/env python
"""Transcribe the reference's own golden vectors into JSON fixtures.

Run HERE (the container that has ``/root/reference``); the GPU box does not
have the reference, so tests only read the committed JSON.  Nothing is
computed in this script: every number is lifted verbatim from the reference's
``test.rs`` files, together with the file:line it came from.

    python tests/golden/make_goldens.py

Outputs (committed):
  tests/golden/conv.json      -- conv{1,2,3}d plain/strided/dilated/grouped + im2col layout
  tests/golden/tensors.json   -- per test-fn ordered tensor literals of the node tests
  tests/golden/tensors_next.json -- the same for the SURVEY.md 8-f nodes (unary family, transpose, mv / vm / vv)
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("NK_REFERENCE", "/root/reference")
NV = os.path.join(REF, "neuronika-variable", "src", "node")
HERE = os.path.dirname(os.path.abspath(__file__))


def fn_blocks(text: str):
    """Yield (name, first_line, body) for every `fn name(...) {` block (brace matched)."""
    for m in re.finditer(r"\bfn\s+(\w+)\s*\([^)]*\)[^{]*\{", text):
        depth, i = 1, m.end()
        while depth and i < len(text):
            c = text[i]
            depth += (c == "{") - (c == "}")
            i += 1
        yield m.group(1), text.count("\n", 0, m.start()) + 1, text[m.end():i - 1]


NUM = r"-?\d+(?:\.\d*)?(?:[eE]-?\d+)?"


def parse_vec(body: str):
    """`vec![a, b, c]` or `vec![v; n]` -> list of floats."""
    body = re.sub(r"//[^\n]*", "", body).strip()
    m = re.fullmatch(rf"\s*({NUM})\s*;\s*(\d+)\s*", body, flags=re.S)
    if m:
        return [float(m.group(1))] * int(m.group(2))
    vals = [v for v in re.split(r"[\s,]+", body) if v]
    return [float(re.sub(r"_?f32$", "", v)) for v in vals]


def find_vecs(body: str):
    """All `vec![...]` literals in order (no nesting in the reference tests)."""
    return [parse_vec(m.group(1)) for m in re.finditer(r"vec!\[([^\]]*)\]", body, flags=re.S)]


def tuple_ints(s: str):
    return [int(v) for v in re.findall(r"\d+", s)]


def gen_conv():
    path = os.path.join(NV, "convolution", "test.rs")
    text = open(path).read()
    cases = {}
    for name, line, body in fn_blocks(text):
        if not re.match(r"(grouped_)?conv[123]d", name):
            continue
        m_in = re.search(r"\(0\.\.([\d_]+)\)", body)
        m_shape = re.search(r"into_shape\(\(([\d,\s]+)\)\)", body)
        m_k = re.search(r"ones\(\(([\d,\s]+)\)\)", body)
        m_s = re.search(r"stride\s*=\s*&\[([\d,\s]+)\]", body)
        m_d = re.search(r"dilation\s*=\s*&\[([\d,\s]+)\]", body)
        m_g = re.search(r"groups\s*=\s*(\d+)", body)
        vecs = {}
        for vm in re.finditer(r"let\s+(true_\w+)\s*(?::[^=]+)?=\s*(?:Array::from_shape_vec\(\s*[^,]+,\s*)?(?:vec|array)!\[(.*?)\]\s*(?:\)\s*\.unwrap\(\))?\s*;", body, flags=re.S):
            vecs[vm.group(1)] = parse_vec(vm.group(2).replace("[", " ").replace("]", " "))
        assert m_in and m_shape and m_k and m_s and m_d, name
        case = {
            "source": f"neuronika-variable/src/node/convolution/test.rs:{line}",
            "input_arange": int(m_in.group(1).replace("_", "")),
            "input_shape": tuple_ints(m_shape.group(1)),
            "kernel_shape": tuple_ints(m_k.group(1)),
            "kernel_fill": 1.0,
            "grad_fill": 1.0,
            "stride": tuple_ints(m_s.group(1)),
            "dilation": tuple_ints(m_d.group(1)),
            "groups": int(m_g.group(1)) if m_g else 1,
        }
        for k in list(vecs):
            if "output" in k:
                case["output"] = vecs[k]
            elif "input_grad" in k:
                case["input_grad"] = vecs[k]
            elif "kernel_grad" in k:
                case["kernel_grad"] = vecs[k]
        assert {"output", "input_grad", "kernel_grad"} <= set(case), (name, list(vecs))
        cases[name] = case
    # im2col layout test
    for name, line, body in fn_blocks(text):
        if name == "im2col":
            arrs = [parse_vec(m.group(1).replace("[", " ").replace("]", " "))
                    for m in re.finditer(r"array!\[(.*?)\];", body, flags=re.S)]
            cases["im2col"] = {
                "source": f"neuronika-variable/src/node/convolution/test.rs:{line}",
                "note": "input = `input` as (3,4,4) stacked twice -> (2,3,4,4); window [1,3,3,3], "
                        "stride 1, dilation 1; expected columns per sample = `expected` reshaped "
                        "(27,4) and transposed -> (4,27)",
                "input": arrs[0], "expected": arrs[1],
            }
    assert len([k for k in cases if "conv" in k]) == 12, sorted(cases)
    return cases


TENSOR_CALL = re.compile(
    r"(new_input|new_backward_input|new_tensor|from_shape_vec)\(\s*\(?([\d,\s]+)\)?\s*,\s*vec!\[([^\]]*)\]",
    flags=re.S)


NEXT_FILES = [   # SURVEY.md 8-f nodes whose (disabled, old-API) tests hold literal vectors
    "negation", "transpose", "power", "sqrt", "sigmoid", "tanh", "softplus", "leaky_relu",
    "matrix_vector_mul", "vector_matrix_mul", "vector_vector_mul",
]


def gen_tensors(files=None):
    files = files or [
        "matrix_matrix_mul", "matrix_matrix_mul_t", "relu", "softmax", "logsoftmax",
        "squared_error", "nll", "sum", "mean", "addition", "pad/zero", "pad/constant",
    ]
    out = {}
    for f in files:
        path = os.path.join(NV, *f.split("/"), "test.rs")
        text = open(path).read()
        per_fn = {}
        for name, line, body in fn_blocks(text):
            tensors = []
            for m in TENSOR_CALL.finditer(body):
                shape = tuple_ints(m.group(2))
                vals = parse_vec(m.group(3))
                n = 1
                for s in shape:
                    n *= s
                if n != len(vals):
                    continue
                tensors.append({"kind": m.group(1), "shape": shape, "values": vals,
                                "line": line + body.count("\n", 0, m.start())})
            scalars = [float(v) for v in re.findall(rf"arr0\(\s*({NUM})\s*\)", body)]
            if tensors or scalars:
                per_fn.setdefault(name, []).append({
                    "source": f"neuronika-variable/src/node/{f}/test.rs:{line}",
                    "tensors": tensors, "scalars": scalars})
        out[f] = per_fn
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: goldens can only be regenerated where the reference is mounted")
    conv = gen_conv()
    with open(os.path.join(HERE, "conv.json"), "w") as fh:
        json.dump(conv, fh, indent=0, separators=(",", ":"))
    tensors = gen_tensors()
    with open(os.path.join(HERE, "tensors.json"), "w") as fh:
        json.dump(tensors, fh, indent=0, separators=(",", ":"))
    nxt = gen_tensors(NEXT_FILES)
    with open(os.path.join(HERE, "tensors_next.json"), "w") as fh:
        json.dump(nxt, fh, indent=0, separators=(",", ":"))
    tensors = dict(tensors, **nxt)
    print("conv cases:", sorted(conv))
    for f, d in tensors.items():
        print(f, {k: [len(b["tensors"]) for b in v] for k, v in d.items()})


if __name__ == "__main__":
    main()
