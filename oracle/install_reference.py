"""Install the UNMODIFIED reference (dsptoolbox v0.8, pure Python) into baseline/_ref/ so that it
travels to the GPU box (baseline/_ref/ is git-ignored but not gpurun-ignored).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: `bench.py --impl reference`, the `cpu_baseline` leg and a few
tests import the reference from there; nothing in dsptoolbox_b200/ does.

Procedure (the recipe of the task contract, then the equivalent for a missing build backend):
  1. `python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
      --target baseline/_ref <copy of /root/reference under /tmp>`
     `--no-deps` because soundfile / sounddevice / matplotlib / seaborn (I/O and plotting only; stubbed by
     oracle/ref_loader.py) are not in the offline wheelhouse.
  2. The reference's build backend is hatchling, which this image does not have (no network to fetch it),
     so step 1 fails with "BackendUnavailable".  A hatchling wheel of this project is nothing but the
     `dsptoolbox/` package directory (pure Python, no generated files: pyproject.toml only points the
     version at dsptoolbox/__init__.py), so the fallback copies exactly that directory -- byte for byte --
     to baseline/_ref/dsptoolbox/, which is what `pip install --target` would have produced.
The outcome is recorded in baseline/_ref/INSTALL.json (method, file count, sha256 over the file list).
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("B2DSP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def _tree_digest(path: str) -> tuple[int, str]:
    h, n = hashlib.sha256(), 0
    for base, _, files in sorted(os.walk(path)):
        if "__pycache__" in base.split(os.sep):
            continue                               # byte-code and numba caches are not part of the package
        for f in sorted(files):
            p = os.path.join(base, f)
            h.update(os.path.relpath(p, path).encode())
            h.update(open(p, "rb").read())
            n += 1
    return n, h.hexdigest()


def install(force: bool = False) -> dict:
    info_path = os.path.join(DST, "INSTALL.json")
    if not os.path.isdir(os.path.join(SRC, "dsptoolbox")):
        return {"status": "skipped", "why": f"{SRC} not present (GPU box): using the prebuilt baseline/_ref"}
    n_src, digest = _tree_digest(os.path.join(SRC, "dsptoolbox"))
    if not force and os.path.exists(info_path):
        old = json.load(open(info_path))
        if old.get("sha256") == digest and os.path.isdir(os.path.join(DST, "dsptoolbox")):
            return old
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    method, pip_tail = "pip", ""
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "__pycache__", "docs", "example_data", "tests"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        pip_tail = (r.stdout + r.stderr)[-400:]
        if r.returncode != 0 or not os.path.isdir(os.path.join(DST, "dsptoolbox")):
            method = "copy of the package directory (pip failed: build backend hatchling unavailable offline)"
            shutil.rmtree(os.path.join(DST, "dsptoolbox"), ignore_errors=True)
            shutil.copytree(os.path.join(SRC, "dsptoolbox"), os.path.join(DST, "dsptoolbox"),
                            ignore=shutil.ignore_patterns("__pycache__"))
    n_dst, digest_dst = _tree_digest(os.path.join(DST, "dsptoolbox"))
    assert digest_dst == digest and n_dst == n_src, "installed tree differs from the reference package"
    info = {"status": "installed", "method": method, "files": n_dst, "sha256": digest, "source": SRC,
            "pip_output_tail": pip_tail}
    json.dump(info, open(info_path, "w"), indent=1)
    return info


if __name__ == "__main__":
    print(json.dumps(install(force="--force" in sys.argv), indent=1))
