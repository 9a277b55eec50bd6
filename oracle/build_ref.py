"""TEST INFRASTRUCTURE -- recipe that builds the reference's own flow model into oracle/_ref/ (git-ignored).

The reference is pure Python, so "building" it means byte-compiling the import closure of
`timewarp.model_constructor.custom_transformer_nvp_constructor` (42 modules: model_constructor.py, model_configs.py,
dataloader.py, modules/**, utilities/**, utils/{chirality,molecule_utils}.py, visualise/visualise.py) from the sources
where they lie under /root/reference into sourceless `.pyc` files, packed into ONE archive oracle/_ref/reference_flow.zip (importable through
zipimport; loose `.pyc` files are dropped by the gpurun snapshot).  No reference SOURCE is copied into the repository; the
archive travels to the GPU box with gpurun like the built `.so`, and `oracle/ref_flow.py` imports it there, so `bench.py --impl reference` times the UNMODIFIED reference modules
(`ConditionalFlowDensityModel.conditional_sample_with_logp` / `.log_likelihood`) on the host cores.

    python -m oracle.build_ref            # no-op when /root/reference is absent (GPU box) or _ref is up to date
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TW_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "reference_flow.zip")
STAMP = os.path.join(OUT_DIR, "reference_flow.python_version")
TOP_FILES = ["__init__.py", "dataloader.py", "model_configs.py", "model_constructor.py"]
TREES = ["modules", "utilities"]
EXTRA = ["utils/__init__.py", "utils/chirality.py", "utils/molecule_utils.py", "visualise/__init__.py", "visualise/visualise.py"]


def closure():
    files = list(TOP_FILES) + list(EXTRA)
    for tree in TREES:
        for d, _, names in os.walk(os.path.join(REF, tree)):
            if "tests" in d.split(os.sep):
                continue
            for n in sorted(names):
                if n.endswith(".py"):
                    files.append(os.path.relpath(os.path.join(d, n), REF))
    return sorted(set(files))


def build_ref(force: bool = False) -> str | None:
    """Returns the archive path, or None when neither the reference tree nor a previously built archive exists."""
    import tempfile
    import warnings
    import zipfile

    if not os.path.isdir(REF):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    ver = sys.version.split()[0]
    files = closure()
    newest = max(os.path.getmtime(os.path.join(REF, rel)) for rel in files)
    if (not force and os.path.exists(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest and os.path.exists(STAMP)
            and open(STAMP).read().strip() == ver):
        return ARCHIVE
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        for rel in files:
            cfile = os.path.join(tmp, rel + "c")
            os.makedirs(os.path.dirname(cfile), exist_ok=True)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", SyntaxWarning)  # the reference's docstrings hold '\\p' escapes
                py_compile.compile(os.path.join(REF, rel), cfile=cfile, dfile=os.path.join("timewarp", rel), doraise=True, optimize=0,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(cfile, os.path.join("timewarp", rel + "c"))  # package `timewarp.*` ...
            if "/" in rel:
                z.write(cfile, rel + "c")  # ... and the top-level packages (`utilities`, `visualise`, ...) it imports besides
    os.replace(ARCHIVE + ".tmp", ARCHIVE)
    with open(STAMP, "w") as f:
        f.write(ver + "\n")
    return ARCHIVE


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
