"""Development builds of the library with extra -D flags: python tools/build_variants.py name=-DFOO,-DBAR=1 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videoyolo_b200 import build
for a in sys.argv[1:]:
    name, flags = a.split("=", 1)
    print(build.build(variant=name, extra_flags=[f for f in flags.split(",") if f]))
