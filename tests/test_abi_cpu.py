"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every declared symbol, and its
host-side geometry entry points reproduce the reference's index tensors.  No GPU compute is called."""
import ctypes
import os
import re

import pytest
import torch

from cliora_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    if _lib.needs_build():
        _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'cliora_b200.h')).read()
    declared = set(re.findall(r'\b(cliora_[a-z0-9_]+)\s*\(', header))
    declared -= {'cliora_status', 'cliora_stream_t'}
    assert declared, 'no declarations parsed'
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared


def test_geometry_matches_reference_indices(lib, golden):
    blob = golden('index.pt')
    from cliora_b200.net.index import Index
    idx = Index()
    for n in range(2, 13):
        assert idx.get_offset(n) == blob[('offset', n)]
        for level in range(1, n):
            l, r = idx.get_inside_index(n, level)
            gl, gr = blob[('inside', n, level)]
            assert torch.equal(l, gl) and torch.equal(r, gr)
        for level in range(0, n - 1):
            p, s = idx.get_outside_index(n, level)
            gp, gs = blob[('outside', n, level)]
            assert torch.equal(p, gp) and torch.equal(s, gs)


def test_layout_and_error_codes(lib):
    lay = _lib.layout(32, 20, 400, 36, True)
    assert lay.rows_in == 32 * 1330 and lay.rows_out == 2 * lay.rows_in and lay.PI == 3
    assert lay.ws_floats > 0 and lay.bws_floats > 0
    d = _lib.Dims(2, 3, 6, 0, 1, 0)      # D % 4 != 0
    out = _lib.Layout()
    assert lib.cliora_chart_layout(ctypes.byref(d), ctypes.byref(out)) == -1
    assert b'bad shape' in lib.cliora_status_string(-1)
    with pytest.raises(_lib.ClioraError):
        _lib.check(-2, 'x')


def test_cpu_tensors_are_rejected_not_computed():
    from cliora_b200.net.diora import DioraMLP
    m = DioraMLP(8)
    with pytest.raises(_lib.ClioraError):
        m(torch.randn(2, 3, 8), None)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under cliora_b200/ (nor bench.py outside its CPU legs) may import
    it, and nothing in the product reads /root/reference at run time."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for dirpath, _, files in os.walk(os.path.join(root, 'cliora_b200')):
        for f in files:
            if not f.endswith('.py'):
                continue
            path = os.path.join(dirpath, f)
            src = open(path).read()
            for node in ast.walk(ast.parse(src)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or '']
                if any(n == 'oracle' or n.startswith('oracle.') for n in names):
                    offenders.append(path)
            if '/root/reference' in src:
                offenders.append(path + ' (reads /root/reference)')
    assert offenders == []


def test_layout_struct_matches_header():
    """The ctypes mirror of cliora_layout lists exactly the int64 fields the header declares, in order."""
    import os
    import re
    from cliora_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, 'include', 'cliora_b200.h')).read()
    body = re.search(r'typedef struct cliora_layout \{(.*?)\} cliora_layout;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in re.findall(r'int64_t\s+([^;]+);', body):
        fields += [x.strip() for x in decl.split(',')]
    assert fields == [name for name, _ in _lib.Layout._fields_]
