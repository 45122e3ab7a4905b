"""A/B of library variants on the small-window regime: python tools/ab_c2.py <lib.so> [...] (runs tools/time_c2.py with each)."""
import os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(2):
    for lib in sys.argv[1:]:
        shutil.copy(lib, os.path.join(root, "gl_slam_b200", "libglba.so"))
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "time_c2.py")], capture_output=True, text=True).stdout
        print(os.path.basename(lib), " | ".join(l for l in out.splitlines() if "solve ms" in l))
