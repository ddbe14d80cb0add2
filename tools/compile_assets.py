"""Compiles the benchmark robots' description files into genesis_b200/assets/compiled/*.json.gz model dumps.

Run in the authoring container (where /root/reference/genesis/assets is mounted):
    python tools/compile_assets.py
The GPU box has no copy of the reference tree; genesis_b200 falls back on these dumps there.
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from genesis_b200.model import parse_mjcf, parse_urdf  # noqa: E402
from genesis_b200.model.compiled import resolve_asset, save_compiled  # noqa: E402

ASSETS = [("xml/franka_emika_panda/panda.xml", "mjcf"), ("urdf/go2/urdf/go2.urdf", "urdf")]

if __name__ == "__main__":
    for file, kind in ASSETS:
        path = resolve_asset(file)
        if path is None:
            raise SystemExit(f"asset not found: {file}")
        l_infos, eqs = (parse_mjcf if kind == "mjcf" else parse_urdf)(path)
        out = save_compiled(file, l_infos, eqs)
        print(f"{file} -> {out} ({os.path.getsize(out) / 1024:.1f} KiB)")
