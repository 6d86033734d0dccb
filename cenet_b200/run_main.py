"""Run one of the reference's scripts UNCHANGED with `networks` = this repo's drop-in package:

    cd /path/to/cenet/src && python -m cenet_b200.run_main main_acdc.py --amp --loss_type boundary ...

`python main_x.py` always puts the script's own directory first on sys.path, where the reference's `networks/` lives; this
launcher keeps that directory importable (the scripts also import `utils`, `datasets`, ...) but places `<repo>/cenet_b200` -- whose
`networks` sub-package re-exports `cenet_b200.networks` -- ahead of it.  Equivalent to `PYTHONPATH=<repo>/cenet_b200 python -P main_x.py`.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m cenet_b200.run_main <script.py> [script arguments ...]", file=sys.stderr)
        return 2
    here = os.path.dirname(os.path.abspath(__file__))
    script = os.path.abspath(argv[0])
    sys.argv = [script] + argv[1:]
    for p in (os.path.dirname(script), here):                   # `here` ends up first
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.modules.pop("networks", None)
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
