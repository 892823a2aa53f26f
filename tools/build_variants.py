"""Build tuning variants of the library next to the default one: python tools/build_variants.py name=-DFLAG=..,-D.. ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rte_b200  # noqa: E402

b = rte_b200.pkg.build
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    out = os.path.join(b.PKG_DIR, f"libore_b200_{name}.so")
    print(b.build_library(force=False, extra_flags=tuple(f for f in flags.split(",") if f), out=out), flush=True)
