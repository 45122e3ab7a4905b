"""Write every BA fixture of tests/golden/*.npz as plain text for tools/ceres_crosscheck.cpp (format in its header).

    python tools/export_scene_text.py [outdir]        # default tests/golden/text (git-ignored; regenerate at will)
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_scene(path, cam, pt, obs_cam, obs_pt, obs_u, obs_v, cam_fixed, K, loss):
    with open(path, "w") as f:
        f.write(f"{len(cam)} {len(pt)} {len(obs_cam)} " + " ".join(repr(float(k)) for k in K) + f" {int(loss)}\n")
        for i in range(len(cam)):
            f.write(f"{int(cam_fixed[i])} " + " ".join(repr(float(x)) for x in cam[i]) + "\n")
        for j in range(len(pt)):
            f.write(" ".join(repr(float(x)) for x in pt[j]) + "\n")
        for k in range(len(obs_cam)):
            f.write(f"{int(obs_cam[k])} {int(obs_pt[k])} {float(obs_u[k])!r} {float(obs_v[k])!r}\n")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "text")
    os.makedirs(out, exist_ok=True)
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        d = np.load(path)
        if "obs_cam" not in d.files:
            continue
        name = os.path.splitext(os.path.basename(path))[0]
        write_scene(os.path.join(out, name + ".txt"), d["cam"], d["pt"], d["obs_cam"], d["obs_pt"], d["obs_u"], d["obs_v"],
                    d["cam_fixed"], d["K"], int(d["loss"]))
        print(name)


if __name__ == "__main__":
    main()
