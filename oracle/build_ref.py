"""TEST INFRASTRUCTURE ONLY -- recipe that "builds" the reference for the GPU box: oracle/_ref/.

The reference is pure Python: there is nothing to compile with gcc, and /root/reference does not
exist on the GPU box.  What can travel is a build artefact: this script byte-compiles the UNMODIFIED reference modules
from where they lie under /root/reference into oracle/_ref/*.pyc (CPython bytecode, a binary; git-ignored, not
gpurun-ignored -- exactly like our own built .so).  No reference source enters the repository.  oracle/ref_harness.py
loads the module from /root/reference when it is mounted and from these .pyc files otherwise, so that bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) time the reference's OWN classes on the box's host cores (`kind: "reference"`).

Run by __graft_entry__.build() whenever /root/reference is present:   python oracle/build_ref.py
"""
import os
import py_compile
import sys

REF_DIR = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = {"llama3.2_model_numpy.py": "llama3_2_model_numpy.pyc", "gemma2_model.py": "gemma2_model.pyc"}


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(REF_DIR):
        if verbose:
            print(f"[build_ref] {REF_DIR} not mounted: keeping whatever is in {OUT}")
        return False
    os.makedirs(OUT, exist_ok=True)
    for src, dst in MODULES.items():
        py_compile.compile(os.path.join(REF_DIR, src), cfile=os.path.join(OUT, dst), dfile=src, doraise=True, optimize=0)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    if verbose:
        print(f"[build_ref] byte-compiled {list(MODULES)} -> {OUT}")
    return True


if __name__ == "__main__":
    build()
