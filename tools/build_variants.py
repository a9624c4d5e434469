"""Tuning helper: build variants of the CUDA library with different launch geometry (used with SB2_LIB=... bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shyft_b200 import _build

VARIANTS = {"pB32x18": ["-DSB2_MINBLOCKS_B=18"], "pB32x24": ["-DSB2_MINBLOCKS_B=24"], "pB16C20": ["-DSB2_MINBLOCKS_B=16", "-DSB2_MINBLOCKS_C=20"], "pB16us32": ["-DSB2_MINBLOCKS_B=16", "-DSB2_UNIT_STEPS=32"],
            "pB32x16": ["-DSB2_MINBLOCKS_B=16"], "pB32x10": ["-DSB2_MINBLOCKS_B=10"], "pB32x8": ["-DSB2_MINBLOCKS_B=8"],
            "us32": ["-DSB2_UNIT_STEPS=32"], "us128": ["-DSB2_UNIT_STEPS=128"], "us256": ["-DSB2_UNIT_STEPS=256"], "us100000": ["-DSB2_UNIT_STEPS=100000"],
            "brent1": ["-DSB2_BRENT_VARIANT=1"], "brent2": ["-DSB2_BRENT_VARIANT=2"],
            "snowfn": ["-DSB2_SNOW_HOT_NOINLINE=1"], "snowfn12": ["-DSB2_SNOW_HOT_NOINLINE=1", "-DSB2_MINBLOCKS_B=12"], "snowfn20": ["-DSB2_SNOW_HOT_NOINLINE=1", "-DSB2_MINBLOCKS_B=20"],
            "pf0": ["-DSB2_PREFETCH_AHEAD=0"], "pf2": ["-DSB2_PREFETCH_AHEAD=2"], "pf8": ["-DSB2_PREFETCH_AHEAD=8"],
            "pB32x12": ["-DSB2_MINBLOCKS_B=12"], "pB32x20": ["-DSB2_MINBLOCKS_B=20"], "pB64x8": ["-DSB2_BLOCK_B=64", "-DSB2_MINBLOCKS_B=8"],
            "pC64x8": ["-DSB2_BLOCK_C=64", "-DSB2_MINBLOCKS_C=8"], "pC32x20": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=20"], "pC32x12": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=12"], "pC64x6": ["-DSB2_BLOCK_C=64", "-DSB2_MINBLOCKS_C=6"], "pC32x24": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=24"], "pC128x4": ["-DSB2_BLOCK_C=128", "-DSB2_MINBLOCKS_C=4"], "pC32x16": ["-DSB2_BLOCK_C=32", "-DSB2_MINBLOCKS_C=16"],
            "nosmemcache": ["-DSB2_CACHE_SMEM=0"], "inl": ["-DSB2_MATH_INLINE=1"], "b128": ["-DSB2_BLOCK=128", "-DSB2_MINBLOCKS=1"], "b64m8": ["-DSB2_BLOCK=64", "-DSB2_MINBLOCKS=8"], "b64m10": ["-DSB2_BLOCK=64", "-DSB2_MINBLOCKS=10"],
            "b32m16": ["-DSB2_BLOCK=32", "-DSB2_MINBLOCKS=16"], "b32m20": ["-DSB2_BLOCK=32", "-DSB2_MINBLOCKS=20"], "b32m14": ["-DSB2_BLOCK=32", "-DSB2_MINBLOCKS=14"], "b32m12": ["-DSB2_BLOCK=32", "-DSB2_MINBLOCKS=12"]}
out_dir = os.path.join(_build.ROOT, "build")
os.makedirs(out_dir, exist_ok=True)
names = sys.argv[1:] or list(VARIANTS)
for name in names:
    _build.build_library(force=True, extra_flags=VARIANTS[name], output=os.path.join(out_dir, f"libshyft_b200_{name}.so"))
    print("built", name)
