"""TEST INFRASTRUCTURE -- recipe that builds the reference's own flow model into oracle/_ref/ (git-ignored).

The reference is pure Python, so "building" it means byte-compiling the import closure of
`timewarp.model_constructor.custom_transformer_nvp_constructor` (42 modules: model_constructor.py, model_configs.py,
dataloader.py, modules/**, utilities/**, utils/{chirality,molecule_utils}.py, visualise/visualise.py) from the sources
where they lie under /root/reference into sourceless `.pyc` files under oracle/_ref/timewarp/.  No reference SOURCE is
copied into the repository; the `.pyc` files travel to the GPU box with gpurun like the built `.so`, and
`oracle/ref_flow.py` imports them there, so `bench.py --impl reference` times the UNMODIFIED reference modules
(`ConditionalFlowDensityModel.conditional_sample_with_logp` / `.log_likelihood`) on the host cores.

    python -m oracle.build_ref            # no-op when /root/reference is absent (GPU box) or _ref is up to date
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TW_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref", "timewarp")
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
    """Returns the output directory, or None when the reference tree is not present (nothing to build from)."""
    if not os.path.isdir(REF):
        return OUT if os.path.isdir(OUT) else None
    stamp = os.path.join(OUT, ".python_version")
    ver = sys.version.split()[0]
    for rel in closure():
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel + "c")  # sourceless layout: module.pyc next to where module.py would be
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src) and os.path.exists(stamp) \
                and open(stamp).read().strip() == ver:
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)  # the reference's docstrings hold '\p' escapes
            py_compile.compile(src, cfile=dst, dfile=os.path.join("timewarp", rel), doraise=True, optimize=0)
    with open(stamp, "w") as f:
        f.write(ver + "\n")
    return OUT


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
