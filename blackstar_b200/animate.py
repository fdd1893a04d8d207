"""``python -m blackstar_b200.animate [-o PATH] [-f] INPUTFILE`` -- mirror of the reference's
`animate` executable (app/Animate.hs): one scene YAML per frame next to each other."""
from __future__ import annotations

import argparse
import os
import sys

from . import animation


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="animate", description="Animation helper for Blackstar")
    ap.add_argument("-o", "--output", default="", metavar="PATH", help="output directory")
    ap.add_argument("-f", "--force", action="store_true", help="overwrite images without asking")
    ap.add_argument("infile", metavar="INPUTFILE")
    a = ap.parse_args(argv)
    if not os.path.isfile(a.infile):
        print("Couldn't open input file.")  # app/Animate.hs:66
        return 0
    try:
        anim = animation.load_animation(a.infile)
    except Exception as e:
        print(f"Error when decoding config:\n{e}")
        return 0
    err = animation.validate_keyframes(anim.keyframes)
    if err:
        print(err)
        return 0
    outdir = a.output or os.getcwd()
    base = os.path.splitext(os.path.basename(a.infile))[0]
    os.makedirs(outdir, exist_ok=True)
    for idx, cfg in enumerate(animation.generate_frames(anim)):
        p = os.path.join(outdir, animation.frame_filename(base, anim.nFrames, idx))
        if os.path.exists(p) and not a.force:
            if input(f"Overwrite {p}? [y/N] ").strip() not in ("y", "Y"):
                print("Nothing was written.")
                continue
        with open(p, "w", encoding="utf-8") as f:
            f.write(animation.config_to_yaml(cfg))
    return 0


if __name__ == "__main__":
    sys.exit(main())
