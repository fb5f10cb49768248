"""Load pbrt_b200/build.py by path and run it.

`import pbrt_b200` deliberately fails when libpbrt_b200.so is missing (there is no fallback), so the
build step cannot go through the package import.  Used by __graft_entry__.build() and the tests.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

ROOT = Path(__file__).resolve().parent


def build_module():
    spec = importlib.util.spec_from_file_location("pbrt_b200_build", ROOT / "pbrt_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(force: bool = False, verbose: bool = False) -> Path:
    return build_module().build(force=force, verbose=verbose)


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
