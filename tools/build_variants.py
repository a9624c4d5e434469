"""Tuning helper: build variants of the CUDA library with different launch geometry (used with SB2_LIB=... bench.py, see tools/tune.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shyft_b200 import _build

VARIANTS = {
    # snow kernel: blocks per SM (register budget), steps per time slice, prefetch distance
    "pB32x12": ["-DSB2_MINBLOCKS_B=12"], "pB32x16": ["-DSB2_MINBLOCKS_B=16"], "pB32x20": ["-DSB2_MINBLOCKS_B=20"],
    "pB64x8": ["-DSB2_BLOCK_B=64", "-DSB2_MINBLOCKS_B=8"], "snowinl": ["-DSB2_SNOW_HOT_NOINLINE=0"],
    # response kernel
    "pC32x12": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=12"], "pC32x20": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=20"],
    "pC32x24": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=24"], "pC64x8": ["-DSB2_BLOCK_C=64", "-DSB2_MINBLOCKS_C=8"],
    "noconst": ["-DSB2_RESP_SMEM_CONST=0"],
    # both
    "us32": ["-DSB2_UNIT_STEPS=32"], "us128": ["-DSB2_UNIT_STEPS=128"], "us256": ["-DSB2_UNIT_STEPS=256"], "nosplit": ["-DSB2_UNIT_STEPS=100000"],
    # interpolation contraction: steps staged per buffer
    "dc16": ["-DSB2_DENSE_TILE_COMPACT=16"], "dc64": ["-DSB2_DENSE_TILE_COMPACT=64"], "btk32": ["-DSB2_DENSE_TILE_BTK=32"], "btk16": ["-DSB2_DENSE_TILE_BTK=16"],
    # snow kernel code footprint (instruction-fetch bound, profiles/README.md)
    "snowflat": ["-DSB2_SNOW_FLAT=1"], "lwcflat": ["-DSB2_LWC_FLAT=1"], "us32b": ["-DSB2_UNIT_STEPS=32"], "us128b": ["-DSB2_UNIT_STEPS=128"],
    "divcall": ["-DSB2_GS_DIV_CALL=1"], "snowcall": ["-DSB2_SNOW_FLAT=0"], "smallsnow": ["-DSB2_GS_DIV_CALL=1", "-DSB2_SNOW_FLAT=0"],
    # hbv step kernel (pt_hs_k / hbv_stack): launch bounds as (block, min blocks per SM)
    "hbvk5": ["-DSB2_HBV_MINBLOCKS_K=5"], "hbvs4": ["-DSB2_HBV_MINBLOCKS_S=4"], "hbvs6": ["-DSB2_HBV_MINBLOCKS_S=6"],
    # forcing-terms kernel: registers (blocks of 128 per SM), steps per thread
    "A8": ["-DSB2_MINBLOCKS_A=8"], "A9": ["-DSB2_MINBLOCKS_A=9"], "A10": ["-DSB2_MINBLOCKS_A=10"], "As4": ["-DSB2_STEPS_A=4"], "As16": ["-DSB2_STEPS_A=16"],
    "hbvus0": ["-DSB2_HBV_UNIT_STEPS=0"], "hbvus32": ["-DSB2_HBV_UNIT_STEPS=32"], "hbvus128": ["-DSB2_HBV_UNIT_STEPS=128"],
    "lwcpair": ["-DSB2_LWC_PAIR=1"],
    "norpB": ["-DSB2_REG_PREFETCH_B=0"], "norpC": ["-DSB2_REG_PREFETCH_C=0"], "norpBC": ["-DSB2_REG_PREFETCH_B=0", "-DSB2_REG_PREFETCH_C=0"],
    "rpA8": ["-DSB2_MINBLOCKS_A=8"], "norpA": ["-DSB2_REG_PREFETCH_A=0"],
    "hps2": ["-DSB2_HPS_MINBLOCKS=2"], "hps3": ["-DSB2_HPS_MINBLOCKS=3"],
    "pf0": ["-DSB2_PREFETCH_AHEAD=0"], "pf2": ["-DSB2_PREFETCH_AHEAD=2"], "pf8": ["-DSB2_PREFETCH_AHEAD=8"],
}
out_dir = os.path.join(_build.ROOT, "build")
os.makedirs(out_dir, exist_ok=True)
names = sys.argv[1:] or list(VARIANTS)
for name in names:
    _build.build_library(force=True, extra_flags=VARIANTS[name], output=os.path.join(out_dir, f"libshyft_b200_{name}.so"))
    print("built", name)
