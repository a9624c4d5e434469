"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/shyft_b200.h declares,
and the product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "shyft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from shyft_b200 import capi
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/shyft_b200.h but not exported by libshyft_b200.so"
    assert set(capi.EXPORTS) <= set(declared)
    assert L.sb2_version() >= 100


def test_library_has_no_python_or_torch_dependency():
    from shyft_b200 import capi
    import subprocess
    out = subprocess.run(["ldd", capi.library_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "libsho_oracle" not in out


def test_struct_layouts_match_the_header():
    from shyft_b200 import capi
    assert capi.GEO_DTYPE.itemsize == 12 * 8
    assert ctypes.sizeof(capi.IdwParameter) == 56
    assert ctypes.sizeof(capi.BtkParameter) == 40
    assert ctypes.sizeof(capi.InterpolationParameter) == 40 + 8 + 5 * 56
    ip = capi.InterpolationParameter()
    assert ip.temperature.sill == 25.0 and ip.temperature.nug == 0.5 and ip.temperature.range == 200000.0 and ip.temperature.zscale == 20.0
    assert ip.temperature_idw.max_members == 20 and ip.precipitation.max_members == 20 and ip.wind_speed.max_members == 10
    assert ip.precipitation.scale_factor == 1.02 and ip.temperature_idw.default_temp_gradient == -0.006 and ip.use_idw_for_temperature == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import shyft_b200
    geo = shyft_b200.geo_cell_data_vector(np.arange(4.0), np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError, match="no CUDA device available"):
        shyft_b200.PTGSKModel(geo)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "shyft_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libsho_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
