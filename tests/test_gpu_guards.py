"""Out-of-bounds writes, checked by the library itself (compute-sanitizer is closed on this GPU pool; profiles/sanitizer_r02.md): with
SB2_GUARD=1 every device buffer is allocated between two 4 KB red zones and sb2_check_guards() verifies them.  The workloads are the ones
tools/sanitizer_case.py holds for memcheck / racecheck: the multi-wave, time-sliced step kernels with TMA-staged dense interpolation, a
windowed run with river routing, the goal kernels (single and batched), and the HBV stacks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = "import sys; sys.path.insert(0, %r); from shyft_b200 import capi; v, msg = capi.check_guards(); print('guards', v, msg); sys.exit(1 if v else 0)" % ROOT


@pytest.mark.parametrize("case,cells,steps", [("ptgsk", 40000, 384), ("routing", 4000, 500), ("goal", 2000, 480)])
def test_red_zones_stay_intact(case, cells, steps):
    env = dict(os.environ, SB2_GUARD="1")
    prog = f"import runpy, sys; sys.argv = ['sanitizer_case.py', {case!r}, '{cells}', '{steps}']; runpy.run_path({os.path.join(ROOT, 'tools', 'sanitizer_case.py')!r}); " + CHECK
    r = subprocess.run([sys.executable, "-c", prog], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert f"{case} ok" in r.stdout and "guards 0" in r.stdout


def test_a_planted_overrun_is_caught():
    """the checker itself: one element written behind the end of a forcing window must be reported"""
    env = dict(os.environ, SB2_GUARD="1")
    prog = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np, torch
import shyft_b200 as sb
from shyft_b200 import capi, synthetic, sharding
geo, ta, env = synthetic.make_region(64, 48, 4, config_index=3)
m = sb.PTGSKOptModel(geo)
m.run_interpolation(sb.InterpolationParameter(), ta, env)
m.set_states(synthetic.default_state(0, 64)); m.run_cells()
assert capi.check_guards()[0] == 0
ptr, rows, cols = m.device_catchment_discharges()
t = torch.as_tensor(sharding.DeviceArrayView(ptr, rows + 1, cols), device='cuda')   # one row more than the buffer has
t[rows, 0] = 1.0
torch.cuda.synchronize()
v, msg = capi.check_guards()
print('guards', v, msg)
sys.exit(0 if v == 1 and 'behind its end' in msg else 1)
"""
    r = subprocess.run([sys.executable, "-c", prog], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
