"""Pre-compiled entity descriptions ("model dumps").

SURVEY.md section 7.1: the step path consumes flat model arrays; robot description files are only needed at
build time.  The GPU box has no copy of the reference's asset tree, so the two robots of the benchmark
configs (Franka Panda MJCF, Unitree Go2 URDF) are compiled once by tools/compile_assets.py with this
package's own parsers and stored here as gzip-JSON (link / joint / geom info with hull vertices).  When the
original asset file is reachable it is parsed instead.
"""

import gzip
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "compiled")

ASSET_DIRS = [p for p in (os.environ.get("GENESIS_B200_ASSETS"), "/root/reference/genesis/assets") if p]


def resolve_asset(file):
    if os.path.isabs(file) and os.path.exists(file):
        return file
    for d in ASSET_DIRS:
        p = os.path.join(d, file)
        if os.path.exists(p):
            return p
    return None


def _key(file, **opts):
    tag = "_".join(f"{k}-{v}" for k, v in sorted(opts.items()) if v is not None)
    base = file.replace("/", "__").replace(".", "_")
    return base + ("__" + tag if tag else "")


def _enc(o):
    if isinstance(o, np.ndarray):
        return {"__nd__": o.tolist(), "dtype": str(o.dtype)}
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, dict):
        return {k: _enc(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_enc(v) for v in o]
    return o


def _dec(o):
    if isinstance(o, dict):
        if "__nd__" in o:
            return np.array(o["__nd__"], dtype=o["dtype"])
        return {k: _dec(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_dec(v) for v in o]
    return o


def save_compiled(file, l_infos, eqs_info, **opts):
    os.makedirs(_DIR, exist_ok=True)
    path = os.path.join(_DIR, _key(file, **opts) + ".json.gz")
    with gzip.open(path, "wt") as f:
        json.dump({"l_infos": _enc(l_infos), "eqs_info": _enc(eqs_info)}, f)
    return path


def load_compiled(file, **opts):
    path = os.path.join(_DIR, _key(file, **opts) + ".json.gz")
    if not os.path.exists(path):
        return None
    with gzip.open(path, "rt") as f:
        d = json.load(f)
    l_infos, eqs = _dec(d["l_infos"]), _dec(d["eqs_info"])
    for li in l_infos:
        for j in li["joints"]:
            for k, v in list(j.items()):
                if isinstance(v, np.ndarray) and v.size == 0:
                    j[k] = v.reshape(0)
    for e in eqs:
        e["objs_name"] = tuple(e["objs_name"])
    return l_infos, eqs
