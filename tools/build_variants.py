#!/usr/bin/env python
"""Build named kernel variants of librast_b200.so for A/B runs on the GPU box.
Usage: python tools/build_variants.py name:-DFOO=1,-DBAR=2 ...   ->  build/variants/librast_b200_<name>.so
Run one with RAST_LIB=build/variants/librast_b200_<name>.so python bench.py ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rasteriser_b200 import build

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defines = [d[2:] if d.startswith("-D") else d for d in defs.split(",") if d]
    print(build.build_variant(name, defines))
