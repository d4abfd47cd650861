"""Installs the UNMODIFIED reference package into ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

    python oracle/install_ref.py            # run in the build container, where /root/reference exists

``oracle/_ref`` is git-ignored (never part of the history) but not gpurun-ignored, so the installed package travels to
the GPU box, where ``/root/reference`` does not exist.  There it serves two purposes, both as the CHECKER or the thing
being compared against, never as the product:

* ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own ``RelGATModel`` (core/model/model.py,
  core/model/layer.py, core/scorer.py, core/loss/*) on the host cores (``kind: "reference"``);
* ``tests/test_gpu_trainer_dropin.py`` runs the reference's own ``RelGATTrainer`` twice — once as shipped on the CPU,
  once with the four class assignments of INTEGRATION.md §1 on the GPU — and compares loss curve and MRR / Hits@k.

The install is the offline recipe of the task contract: ``pip install --no-index --no-build-isolation --no-deps
--target oracle/_ref <copy of /root/reference>`` (the copy under /tmp because /root/reference is read-only and
setuptools writes build/ and *.egg-info into the source tree).  ``torch_scatter`` is NOT part of the reference
(third-party, absent from this image): ``oracle/standin/torch_scatter`` restates its two functions; the two
logging-only packages the reference imports are stubbed by ``oracle/ref_shim.py``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("RELGAT_REFERENCE_ROOT", "/root/reference")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "relgat_projector", "core", "model", "layer.py"))


def install(force: bool = False) -> str:
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "relgat_projector")):
        raise RuntimeError(f"reference not found under {SOURCE}")
    tmp = tempfile.mkdtemp(prefix="relgat_ref_")
    try:
        copy = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, copy)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, copy]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout.decode()[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
