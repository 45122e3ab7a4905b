"""A/B kernel timings for library variants: python tools/ab_kernels.py <lib.so> [...]; prints per-kernel ms on C4."""
import json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    shutil.copy(lib, os.path.join(root, "gl_slam_b200", "libglba.so"))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--lm-iters", "0"],
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(os.path.basename(lib), "step ms %.4f  value %.3e" % (d["ms_per_step"], d["value"]),
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(lib, "FAILED", e, out.stderr[-500:])
