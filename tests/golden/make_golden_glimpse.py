"""
Golden vectors for ``bin_hist`` from the REFERENCE's own source (run in the build container, where
/root/reference exists):  python tests/golden/make_golden_glimpse.py

glimpse_reader.py imports matplotlib (absent here), so the function is taken out of the file's syntax tree and
executed on its own -- unmodified -- with torch / typing in scope.  Output: tests/golden/ref_glimpse.pt.
"""

import ast
from pathlib import Path
from typing import Tuple  # noqa: F401  (used by the extracted function's annotations)

import torch

REF = Path("/root/reference/tapqir/imscroll/glimpse_reader.py")
OUT = Path(__file__).resolve().parent / "ref_glimpse.pt"


def reference_bin_hist():
    tree = ast.parse(REF.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "bin_hist")
    scope = {"torch": torch, "Tuple": Tuple}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF), "exec"), scope)
    return scope["bin_hist"]


def main():
    bin_hist = reference_bin_hist()
    g = torch.Generator().manual_seed(0)
    cases = []
    for n, s in [(1, 3), (2, 1), (7, 3), (10, 3), (11, 3), (33, 4), (101, 7), (64, 1), (5, 9)]:
        samples = torch.sort(torch.randperm(400, generator=g)[:n] + 50).values.to(torch.int)
        weights = torch.rand(n, generator=g, dtype=torch.float64)
        weights = weights / weights.sum()
        for default in (torch.float32, torch.float64):
            torch.set_default_dtype(default)
            out_s, out_w = bin_hist(samples, weights, s)
            cases.append(dict(samples=samples, weights=weights, s=s, default=str(default), out_samples=out_s, out_weights=out_w))
    torch.set_default_dtype(torch.float32)
    torch.save(cases, OUT)
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    main()
