"""Developer tool: build tuning variants of the library (extra -D switches) into build_variants/.

    python scripts/build_variants.py name1="-DPTOY_TILE_CTAS=5" name2="-DPTOY_TILE=128 -DPTOY_TILE_CTAS=8"
Then on the GPU box:  PTOY_B200_LIB=build_variants/libptoy_name1.so python scripts/time_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptoy_b200 import build  # noqa: E402

os.makedirs(os.path.join(ROOT, "build_variants"), exist_ok=True)
for arg in sys.argv[1:]:
    name, flags = arg.split("=", 1)
    lib = os.path.join(ROOT, "build_variants", f"libptoy_{name}.so")
    build.build(force=True, lib=lib, extra=flags.split())
    print(lib)
