"""Builds experiment variants of the library (compile-time switches) next to the production one:
    python tools/build_variants.py name=DEF1,DEF2 name2=DEF3 ...   ->  pointwise_b200/lib/variants/<name>.so
Run one with CONV3P_LIB=pointwise_b200/lib/variants/<name>.so (tools/ab_variants.sh times them in one GPU call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointwise_b200 import build as b  # noqa: E402

vdir = os.path.join(b.LIB_DIR, "variants")
os.makedirs(vdir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(vdir, name + ".so")
    print(b.build(force=True, defines=[d for d in defs.split(",") if d], out=out))
