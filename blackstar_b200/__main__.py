"""``python -m blackstar_b200`` -- host-side mirror of the ``blackstar`` executable
(app/Main.hs): same flags, same batch-directory behaviour, same file naming; the render,
supersample, bloom and sRGB/8-bit map run on the GPU through the C ABI.

    blackstar [-p|--preview] [-o|--output PATH] [-f|--force] [-s|--starmap PATH] INPUTFILE

Differences, all forced by what this image lacks: the star map is a PPM-format binary
catalogue (what ``generate-tree`` reads, src/StarMap.hs:45-58), not a cereal-encoded
``stars.kdt``; ``--starmap synthetic`` uses the deterministic synthetic catalogue.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

from . import config, starmap


def _prompt_overwrite(path: str) -> bool:
    """Util.promptOverwriteFile (src/Util.hs:18-27)."""
    if not os.path.exists(path):
        return True
    ans = input(f"Overwrite {path}? [y/N] ")
    return ans.strip().lower() == "y"


def handle_scene(r, args, outdir: str, filename: str):
    """Main.handleScene + doRender (app/Main.hs:80-125)."""
    from .render import write_img
    name = os.path.splitext(os.path.basename(filename))[0]
    print(f"Reading {filename}...")
    try:
        cfg = config.load_config(filename)
    except Exception as e:  # prettyPrintParseException, then carry on (app/Main.hs:91)
        print(e)
        return
    print("Scene successfully read.")
    if args.preview:
        name = "prev-" + name
    cfg = config.prepare_scene(cfg, args.preview)
    print(f"Rendering {name}...")
    t = time.perf_counter()
    if cfg.scene.bloomStrength != 0:
        print("Applying bloom...")
    img8 = r.do_render_srgb8(cfg)
    st = r.last_stats
    print(f"Rendering completed in {time.perf_counter() - t:.3f} seconds "
          f"({st['rays'] / max(st['trace_ms'], 1e-9) / 1e3:.1f} Mrays/s in the trace kernel).")
    out_name = os.path.join(outdir, name + ".png")
    print(f"Saving to {out_name}...")
    if args.force or _prompt_overwrite(out_name):
        write_img(img8, out_name)
    print("Everything done. Thank you!")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="blackstar", description="Blackstar v0.1 (B200)")
    ap.add_argument("-p", "--preview", action="store_true", help="preview render (small size)")
    ap.add_argument("-o", "--output", default="", metavar="PATH", help="output directory")
    ap.add_argument("-f", "--force", action="store_true", help="overwrite images without asking")
    ap.add_argument("-s", "--starmap", default="stars.kdt", metavar="PATH", help="path to starmap")
    ap.add_argument("inputfile", metavar="INPUTFILE")
    args = ap.parse_args(argv)

    from .render import Renderer
    try:
        r = Renderer(n_gpus=0)
    except Exception as e:
        print(e)
        return 1
    try:
        if args.starmap == "synthetic":
            r.set_stars(starmap.synthetic_stars())
        else:
            with open(args.starmap, "rb") as f:
                r.set_stars_file(f.read())   # stars.kdt (the reference's tree file) or the PPM catalogue
    except Exception as e:  # app/Main.hs:50
        print(f"Error decoding star tree: \n{e}")
        return 1
    print("Starmap successfully read.")
    outdir = os.path.abspath(args.output or os.getcwd())
    os.makedirs(outdir, exist_ok=True)
    filename = os.path.abspath(args.inputfile)
    if os.path.isdir(filename):  # app/Main.hs:64-77
        print(f"{filename} is a directory. Rendering all scenes inside it...")
        files = sorted(f for f in os.listdir(filename) if os.path.splitext(f)[1] == ".yaml")
        for i, f in enumerate(files, 1):
            print(f"Batch mode progress: {i}/{len(files)}")
            handle_scene(r, args, outdir, os.path.join(filename, f))
    else:
        handle_scene(r, args, outdir, filename)
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
