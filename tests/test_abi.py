"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls:
there is no GPU in the CPU test tier), and the product has no route into oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    inc = os.path.join(ROOT, "include")
    src = "".join(open(os.path.join(inc, f)).read() for f in sorted(os.listdir(inc)) if f.endswith(".h"))
    return sorted(set(re.findall(r"MAGI_API\s+[\w\s\*]+?\b(magi_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from magi_v2_b200 import _lib
    names = _declared()
    assert len(names) >= 12
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    L = ctypes.CDLL(_lib.LIB_PATH)
    for nm in names:
        assert hasattr(L, nm), nm
    assert _lib.lib().magi_b200_abi_version() == _lib.ABI_VERSION == 3


def test_host_only_entry_points():
    from magi_v2_b200 import _lib
    L = _lib.lib()
    D, P = ctypes.c_int(), ctypes.c_int()
    for name, (d, p) in {"seir3": (3, 3), "seir4": (4, 3), "sirw": (4, 5), "lorenz96": (10, 1)}.items():
        assert L.magi_b200_model_dims(_lib.MODEL_IDS[name], ctypes.byref(D), ctypes.byref(P)) == 0
        assert (D.value, P.value) == (d, p)
    assert L.magi_b200_model_dims(99, ctypes.byref(D), ctypes.byref(P)) == -1
    assert L.magi_b200_packed_bytes(2, 4, 161) == 2 * 4 * 3 * 168 * 168 * 8
    assert L.magi_b200_status_string(0) == b"ok"
    assert L.magi_b200_factor_workspace_bytes(10, 161) == 10 * (3 * 161 * 161 + 56 * 161) * 8   # 56 = block size of factor.cu
    # argument validation happens before any CUDA call
    assert L.magi_b200_cov_build(None, 0, None, None, 2.01, 1, 1, 5, 0, None, None, None, None) == -1
    assert L.magi_b200_pack_matrices(None, None, None, 1, 1, 5, None, None) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from magi_v2_b200 import ops
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.cov_build(torch.zeros(5, dtype=torch.float64), torch.ones(1, 1, dtype=torch.float64),
                      torch.ones(1, 1, dtype=torch.float64), 2.01, False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "magi_v2_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
    code = "import sys; import magi_v2_b200, magi_v2_b200.ops, magi_v2_b200.magi; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
