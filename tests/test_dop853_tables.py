"""include/rb_dop853_tables.h is GENERATED (tools/gen_dop853_tables.py) from the DOP853 coefficients of the scipy installed
here -- the third-party dependency the reference hands `sim_type='DOP853'` to (spacecraft_model.py:516).  The committed header
must be what the generator writes, and its numbers must be scipy's, bit for bit."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rb_dop853_tables.h")


def _rows(text, name):
    body = re.search(r"#define %s (.*?)(?=\n#define|\n#endif)" % name, text, re.S).group(1).replace("\\\n", " ")
    return [float.fromhex(t) if t != "0" else 0.0 for t in re.findall(r"-?0x[0-9a-f.]+p[+-]?\d+|(?<![\w.])0(?![\w.])", body)]


def test_header_holds_scipys_coefficients_bit_for_bit():
    D = pytest.importorskip("scipy.integrate._ivp.dop853_coefficients")
    text = open(HEADER).read()
    assert np.array_equal(np.array(_rows(text, "RB_DOP_C_INIT")), D.C)
    assert np.array_equal(np.array(_rows(text, "RB_DOP_A_INIT")).reshape(16, 15), D.A[:, :15])
    assert np.all(D.A[:, 15] == 0)                                   # (the column the header leaves out)
    assert np.array_equal(np.array(_rows(text, "RB_DOP_E3_INIT")), D.E3)
    assert np.array_equal(np.array(_rows(text, "RB_DOP_E5_INIT")), D.E5)
    assert np.array_equal(np.array(_rows(text, "RB_DOP_D_INIT")).reshape(4, 16), D.D)
    assert np.array_equal(D.B, D.A[12, :12])                         # row 12 of A is B (what both steppers use)


def test_generator_reproduces_the_committed_header(tmp_path):
    pytest.importorskip("scipy")
    out = str(tmp_path / "rb_dop853_tables.h")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_dop853_tables.py"), out], check=True, capture_output=True)
    assert open(out).read() == open(HEADER).read()
