#!/usr/bin/env python
"""Package the UNMODIFIED reference (dfm/emcee, pure Python) for the CPU arm of bench.py.

    python baseline/make_ref.py        # authoring container only: needs /root/reference

Writes ``baseline/_ref/emcee_reference.zip`` (git-ignored, but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot): the files of ``/root/reference/src/emcee``
byte for byte, plus the one-line ``emcee/emcee_version.py`` that setuptools_scm would
generate at install time (``src/emcee/__init__.py:22`` imports it; ``setup.py:59-64``).
``bench.py --impl reference`` puts the archive on ``sys.path`` (zipimport) and drives the
reference's own ``EnsembleSampler`` -- nothing of this repository is on that path.
The reference cannot be pip-installed here (its build backend needs setuptools_scm and
network access); copying the package directory is what ``pip install`` would do.
"""
import os
import sys
import zipfile

REF = "/root/reference/src/emcee"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "emcee_reference.zip")


def main():
    if not os.path.isdir(REF):
        print("make_ref: %s not present (GPU box?) -- keeping %s" % (REF, OUT))
        return 0 if os.path.exists(OUT) else 1
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    n = 0
    with zipfile.ZipFile(OUT, "w", zipfile.ZIP_DEFLATED) as z:
        for root, _dirs, files in os.walk(REF):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                full = os.path.join(root, f)
                z.write(full, os.path.join("emcee", os.path.relpath(full, REF)))
                n += 1
        z.writestr("emcee/emcee_version.py", '__version__ = "3.1.6+reference.8ab6c0f"\n')
    print("make_ref: %d files -> %s" % (n + 1, OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
