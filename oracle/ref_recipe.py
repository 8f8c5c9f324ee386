"""Recipe that makes the pure-Python reference (quantik-core 1.1.0) available next to the oracle.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rule as the rest of ``oracle/``): only ``tests/``,
``__graft_entry__`` and ``bench.py``'s CPU legs use it; the product package never does.

``populate()`` runs in the BUILD container, where ``/root/reference`` exists: it copies the reference's
own package (``src/quantik_core/**/*.py``) and its measuring harness (``benchmarks/*.py``) -- unmodified --
into the git-ignored ``oracle/_ref/`` and writes a manifest with one sha256 per file.  ``oracle/_ref/``
is NOT in ``.gpurunignore``, so it travels to the GPU box like the built ``.so`` files do; nothing on the
box ever reads ``/root/reference``.  The repo's history holds this recipe only, never reference sources.

    python -m oracle.ref_recipe            # populate oracle/_ref/ (no-op without /root/reference)
    python -m oracle.ref_recipe --time 10  # time the reference's own rollout loop on all host cores
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import time
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("QUANTIK_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(_HERE, "_ref")
MANIFEST = os.path.join(REF_DST, "MANIFEST.json")


def _copy_py_tree(src: str, dst: str, recursive: bool, files: dict) -> None:
    for root, dirs, names in os.walk(src):
        if not recursive:
            dirs[:] = []
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, src)
        for name in sorted(names):
            if not (name.endswith(".py") or name == "py.typed"):
                continue
            out_dir = os.path.normpath(os.path.join(dst, rel))
            os.makedirs(out_dir, exist_ok=True)
            shutil.copyfile(os.path.join(root, name), os.path.join(out_dir, name))
            with open(os.path.join(out_dir, name), "rb") as f:
                files[os.path.relpath(os.path.join(out_dir, name), REF_DST)] = hashlib.sha256(f.read()).hexdigest()


def populate(force: bool = False) -> Optional[str]:
    """Copy the reference package into oracle/_ref/ (build container only).  -> the path, or None when
    neither the reference nor an earlier copy exists."""
    src_pkg = os.path.join(REF_SRC, "src", "quantik_core")
    if not os.path.isdir(src_pkg):
        return REF_DST if os.path.exists(MANIFEST) else None
    if os.path.exists(MANIFEST) and not force:
        newest = max(os.path.getmtime(os.path.join(r, n)) for r, _, ns in os.walk(src_pkg) for n in ns)
        if os.path.getmtime(MANIFEST) >= newest:
            return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    files: dict = {}
    _copy_py_tree(src_pkg, os.path.join(REF_DST, "quantik_core"), True, files)
    _copy_py_tree(os.path.join(REF_SRC, "benchmarks"), os.path.join(REF_DST, "benchmarks"), False, files)
    version = "unknown"
    try:
        with open(os.path.join(REF_SRC, "pyproject.toml")) as f:
            for line in f:
                if line.strip().startswith("version"):
                    version = line.split("=", 1)[1].strip().strip('"')
                    break
    except OSError:
        pass
    with open(MANIFEST, "w") as f:
        json.dump({"reference": "mberlanda/quantik-core-py", "version": version, "copied_from": REF_SRC,
                   "note": "unmodified copies, build output, git-ignored", "files": files}, f, indent=1, sort_keys=True)
    return REF_DST


def available() -> bool:
    return os.path.exists(MANIFEST)


def path() -> str:
    """sys.path entry under which ``import quantik_core`` / ``import benchmarks`` resolve to the reference."""
    if not available():
        raise RuntimeError("oracle/_ref/ is missing: run __graft_entry__.build() (or python -m oracle.ref_recipe) "
                           "in the build container, where /root/reference exists")
    return REF_DST


def import_reference():
    """-> the reference's ``quantik_core`` package (from oracle/_ref/)."""
    p = path()
    if p not in sys.path:
        sys.path.insert(0, p)
    import quantik_core
    return quantik_core


def verify() -> bool:
    """Every file of oracle/_ref/ still has the sha256 the manifest recorded (i.e. is the unmodified reference)."""
    with open(MANIFEST) as f:
        m = json.load(f)
    for rel, digest in m["files"].items():
        with open(os.path.join(REF_DST, rel), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                return False
    return True


# ----------------------------------------------------------------------------- timing the reference itself
def _rollout_worker(args):
    """One process: the reference's own BeamSearchEngine._rollout (beam_search.py:624-645) from the empty
    board, back to back for `seconds`.  -> (playouts, P0 wins, elapsed)."""
    seconds, seed, ref_path = args
    if ref_path not in sys.path:
        sys.path.insert(0, ref_path)
    from quantik_core.beam_search import BeamSearchConfig, BeamSearchEngine
    eng = BeamSearchEngine(BeamSearchConfig(random_seed=seed))
    empty = (0, 0, 0, 0, 0, 0, 0, 0)
    n = p0 = 0
    t0 = time.perf_counter()
    while True:
        for _ in range(8):
            p0 += eng._rollout(empty) > 0
        n += 8
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n, p0, dt


def time_reference_playouts(seconds: float = 10.0, procs: Optional[int] = None, seed: int = 20260712):
    """The reference's rollout loop on `procs` processes (default: every core of the affinity mask),
    independent seeds.  -> dict(value playouts/s, cores, playouts, p0_rate, seconds)."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    procs = procs or len(os.sched_getaffinity(0))
    ref = path()
    if procs == 1:
        res = [_rollout_worker((seconds, seed, ref))]
    else:
        with ProcessPoolExecutor(procs, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_rollout_worker, [(seconds, seed + i, ref) for i in range(procs)]))
    n = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return {"value": n / wall, "cores": procs, "playouts": n, "p0_rate": sum(r[1] for r in res) / n, "seconds": wall}


if __name__ == "__main__":
    out = populate(force="--force" in sys.argv)
    print("oracle/_ref:", out, "verified" if out and verify() else "")
    if "--time" in sys.argv:
        secs = float(sys.argv[sys.argv.index("--time") + 1])
        print(json.dumps(time_reference_playouts(secs, 1)))
        print(json.dumps(time_reference_playouts(secs)))
