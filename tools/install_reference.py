#!/usr/bin/env python
"""Install the UNMODIFIED reference (cournape/arnoldi-py) into the git-ignored baseline/_ref/.

    python tools/install_reference.py [/root/reference]

The contract's recipe is
    pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
        --target baseline/_ref /root/reference
but the reference's build backend is hatchling (pyproject.toml:17-19), which is in neither this
image nor the offline wheelhouse, so that command dies with "No module named 'hatchling'".  The
package is pure Python, so the same pip install is run on a scratch copy under /tmp whose
pyproject.toml names setuptools as the backend instead (packaging metadata only -- no file under
src/arnoldi is touched; the installed modules are byte-identical to the reference's, which this
script verifies).  baseline/_ref/ is git-ignored (never committed) but travels to the GPU box.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "arnoldi"
version = "{version}"
requires-python = ">=3.11"

[tool.setuptools.packages.find]
where = ["src"]
"""


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(ref, "src", "arnoldi")
    if not os.path.isdir(src):
        print(f"install_reference: {src} not found; nothing installed")
        return 1
    ns = {}
    with open(os.path.join(src, "_version.py")) as f:
        exec(f.read(), ns)
    with tempfile.TemporaryDirectory(prefix="arnoldi_ref_") as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(ref, work, ignore=shutil.ignore_patterns(".git", ".venv", "__pycache__"))
        with open(os.path.join(work, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT.format(version=ns.get("__version__", "0.0.0")))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, work]
        subprocess.run(cmd, check=True)
    inst = os.path.join(TARGET, "arnoldi")
    names = sorted(n for n in os.listdir(src) if n.endswith(".py"))
    match, mismatch, errors = filecmp.cmpfiles(src, inst, names, shallow=False)
    assert not mismatch and not errors and match == names, (mismatch, errors)
    print(f"install_reference: {len(names)} modules installed into {TARGET}, byte-identical to {src}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
