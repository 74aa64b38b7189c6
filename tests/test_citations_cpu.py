"""Every reference citation `path/file.jl:line[-line]` in the boundary header and the design docs points at an existing
file of the reference with at least that many lines (skipped where /root/reference is not mounted, e.g. on the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DOCS = ["include/lsob200.h", "DESIGN.md", "INTEGRATION.md", "oracle/reference_port.py", "leastsquaresoptim.jl_b200/api.py"]
PAT = re.compile(r"((?:src|test|benchmark)/[\w/]+\.jl):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("doc", DOCS)
def test_cited_reference_lines_exist(doc):
    text = open(os.path.join(ROOT, doc), encoding="utf-8").read()
    cites = PAT.findall(text)
    assert cites, f"{doc} cites no reference lines"
    nlines = {}
    for path, a, b in cites:
        full = os.path.join(REF, path)
        assert os.path.isfile(full), f"{doc}: cited file {path} does not exist in the reference"
        if path not in nlines:
            nlines[path] = sum(1 for _ in open(full, encoding="utf-8", errors="replace"))
        last = int(b) if b else int(a)
        assert 1 <= int(a) <= last <= nlines[path], f"{doc}: {path}:{a}{'-' + b if b else ''} is outside the file ({nlines[path]} lines)"
