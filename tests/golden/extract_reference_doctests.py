#!/usr/bin/env python
"""Extracts every doctest vector of the hot path from the reference's sources into
tests/golden/reference_doctests.json.

Run in the build container, where /root/reference exists (the GPU box has no copy of the
reference: the tests read only the committed JSON):

    python tests/golden/extract_reference_doctests.py [/root/reference]

Each record: file, line (of the `#Nx.Tensor<` that prints the value), the iex> expressions of
the example up to that point, the printed type, shape and values (complex numbers as [re, im]).
Nothing is interpreted: the expressions are kept as text so that a reader can see what each
vector pins; tests/test_golden_fixtures.py maps (file, line) to oracle / GPU calls.
"""
import json
import os
import re
import sys

FILES = ["lib/nx_signal.ex", "lib/nx_signal/windows.ex", "lib/nx_signal/convolution.ex", "lib/nx_signal/filters.ex",
         "lib/nx_signal/transforms.ex", "lib/nx_signal/waveforms.ex"]

NUM = r"[-+]?(?:NaN|Inf|\d+(?:\.\d+)?(?:e[-+]?\d+)?)"
CPX = re.compile(rf"({NUM})([-+](?:NaN|Inf|\d+(?:\.\d+)?(?:e[-+]?\d+)?))i")


def parse_values(text):
    """Nx's printed nested list -> Python nested list (complex a+bi -> [a, b])."""
    text = CPX.sub(lambda m: f"[{m.group(1)}, {m.group(2)}]", text)
    text = re.sub(r"\bNaN\b", "float('nan')", text)
    text = re.sub(r"(?<![\w.])-Inf\b", "float('-inf')", text)
    text = re.sub(r"\bInf\b", "float('inf')", text)
    return eval(text, {"__builtins__": {}, "float": float})  # numbers and brackets only


def extract(path, rel):
    lines = open(path).read().split("\n")
    out, exprs, i = [], [], 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln.startswith("iex>") or ln.startswith("...>"):
            if ln.startswith("iex>"):
                exprs.append(ln[4:].strip())
            else:
                exprs[-1] += " " + ln[4:].strip()
            i += 1
            continue
        if ln.startswith("#Nx.Tensor<"):
            j = i + 1
            block = []
            while lines[j].strip() != ">":
                block.append(lines[j].strip())
                j += 1
            vec = []
            if block[0].startswith("vectorized"):
                vec = [int(re.sub(r".*:\s*", "", d)) for d in re.findall(r"\[([^\]]*)\]", block[0])]
                block = block[1:]
            m = re.match(r"(\w+)((?:\[[^\]]*\])*)$", block[0])
            typ = m.group(1)
            shape = [int(re.sub(r".*:\s*", "", d)) for d in re.findall(r"\[([^\]]*)\]", m.group(2))]
            values = parse_values(" ".join(block[1:]))
            out.append({"file": rel, "line": i + 1, "exprs": list(exprs), "type": typ, "vectorized": vec,
                        "shape": shape, "values": values})
            i = j + 1
            continue
        exprs = []  # a blank line or prose ends the example
        i += 1
    return out


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    records = []
    for rel in FILES:
        records += extract(os.path.join(root, rel), rel)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_doctests.json"), "w") as f:
        json.dump({"reference": "elixir-nx/nx_signal v0.3.0 @ dcf5b81", "records": records}, f, indent=0)
    print(f"{len(records)} doctest vectors from {len(FILES)} files")


if __name__ == "__main__":
    main()
