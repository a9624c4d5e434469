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


def test_ctypes_structs_agree_with_the_c_compiler(tmp_path):
    """sizeof / offsetof of every struct that crosses the ABI, as gcc lays them out from include/shyft_b200.h, against the ctypes mirrors"""
    import shutil
    import subprocess
    from shyft_b200 import capi
    from shyft_b200.state_io import ID_DTYPE
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "shyft_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(sb2_geo_cell), sizeof(sb2_idw_parameter), sizeof(sb2_btk_parameter), sizeof(sb2_interpolation_parameter),
           sizeof(sb2_target), sizeof(sb2_q_adjust_result), sizeof(sb2_cell_state_id));
    printf("%zu %zu %zu %zu %zu\n", offsetof(sb2_target, n_catchments), offsetof(sb2_target, scale_factor), offsetof(sb2_target, property),
           offsetof(sb2_target, period_points_us), offsetof(sb2_q_adjust_result, diagnostics));
    return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.check_call([gcc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes, offsets = [list(map(int, line.split())) for line in subprocess.check_output([str(exe)], text=True).splitlines()]
    assert sizes == [capi.GEO_DTYPE.itemsize, ctypes.sizeof(capi.IdwParameter), ctypes.sizeof(capi.BtkParameter),
                     ctypes.sizeof(capi.InterpolationParameter), ctypes.sizeof(capi.Target), ctypes.sizeof(capi.QAdjustResult), ID_DTYPE.itemsize]
    assert offsets == [capi.Target.n_catchments.offset, capi.Target.scale_factor.offset, capi.Target.property.offset,
                       capi.Target.period_points_us.offset, capi.QAdjustResult._diagnostics.offset]
