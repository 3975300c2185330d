"""The lookup tables carried by the product (include/monty_b200/detail/monty_tables.inc)
and by the C oracle are DATA extracted from the reference
(tools/extract_tables.py).  Check them against their mathematical definition
(ref: scripts/update_ziggurat_tables, scripts/update_gamma_tables) and, when the
compiled reference is present, bit for bit."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from oracle_lib import ORACLE_DIR, ROOT, have_reference


def parse_tables(path):
    txt = open(path).read()
    out = {}
    for name in ("MONTY_ZIG_X_VALUES", "MONTY_ZIG_Y_VALUES", "MONTY_TAIL_F_VALUES", "MONTY_TAIL_D_VALUES"):
        m = re.search(r"#define %s \\\n((?:.*\\\n)*.*)\n" % name, txt)
        body = m.group(1).replace("\\", " ")
        vals = [float.fromhex(v.strip().rstrip("f")) for v in body.split(",") if v.strip()]
        out[name] = np.array(vals)
    return out


@pytest.fixture(scope="module")
def tables():
    return parse_tables(os.path.join(ROOT, "include", "monty_b200", "detail", "monty_tables.inc"))


def test_product_and_oracle_tables_identical():
    a = open(os.path.join(ROOT, "include", "monty_b200", "detail", "monty_tables.inc")).read()
    b = open(os.path.join(ROOT, "oracle", "oracle_tables.inc")).read()
    assert a == b


def test_ziggurat_tables_definition(tables):
    x, y = tables["MONTY_ZIG_X_VALUES"], tables["MONTY_ZIG_Y_VALUES"]
    assert x.size == 257 and y.size == 256
    assert x[256] == 0.0 and (np.diff(x) < 0).all()
    # y[i] = x[i+1] / x[i]   (scripts/update_ziggurat_tables: y <- x[-1] / x[seq_len(n)])
    assert np.allclose(y, x[1:] / x[:-1], rtol=1e-15, atol=0)
    # equal-area layers: x[i] * (f(x[i+1]) - f(x[i])) == v for 1 <= i < 255, v = x[0] * f(x[1])
    f = lambda t: np.exp(-t * t / 2)  # noqa: E731
    v = x[0] * f(x[1])
    areas = x[1:255] * (f(x[2:256]) - f(x[1:255]))
    assert np.allclose(areas, v, rtol=1e-8)
    # and the base layer: r * f(r) + tail integral
    r = x[1]
    tail = math.sqrt(math.pi / 2) * math.erfc(r / math.sqrt(2))
    assert abs((r * f(r) + tail) - v) / v < 1e-8


def test_stirling_tail_definition(tables):
    tf, td = tables["MONTY_TAIL_F_VALUES"], tables["MONTY_TAIL_D_VALUES"]
    assert tf.size == 128 and td.size == 10
    k = np.arange(128, dtype=np.float64)
    ref = np.array([math.lgamma(i + 1) for i in k]) - (math.log(math.sqrt(2 * math.pi)) + (k + 0.5) * np.log(k + 1) - (k + 1))
    assert np.allclose(td, ref[:10], rtol=1e-10)
    assert np.allclose(tf, ref.astype(np.float32), rtol=5e-4, atol=1e-8)
    assert np.array_equal(tf[:10], td.astype(np.float32))


def test_tables_equal_reference(tables):
    if not have_reference():
        pytest.skip("compiled reference not present")
    lib = ctypes.CDLL(os.path.join(ORACLE_DIR, "_ref", "libmonty_ref.so"))
    zx = (ctypes.c_double * 257)(); zy = (ctypes.c_double * 256)()
    tf = (ctypes.c_float * 128)(); td = (ctypes.c_double * 10)()
    lib.ref_tables(zx, zy, tf, td)
    assert np.array_equal(np.array(zx), tables["MONTY_ZIG_X_VALUES"])
    assert np.array_equal(np.array(zy), tables["MONTY_ZIG_Y_VALUES"])
    assert np.array_equal(np.array(tf, dtype=np.float64), tables["MONTY_TAIL_F_VALUES"])
    assert np.array_equal(np.array(td), tables["MONTY_TAIL_D_VALUES"])
