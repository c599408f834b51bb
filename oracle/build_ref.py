"""ORACLE (test infrastructure) recipe: vendor the reference's own Python package into oracle/_ref/ so that the CPU arm
of bench.py (`--impl reference`, `cpu_baseline`) times the REFERENCE'S code (fam.llm.fast_model.Transformer +
fam.llm.fast_inference_utils.generate) on the GPU box, where /root/reference does not exist.

    python oracle/build_ref.py          (run by __graft_entry__.build() whenever /root/reference is mounted)

Nothing is modified: the files are copied byte for byte from where they lie under /root/reference (``fam/**/*.py``
only -- the package is pure Python, there is nothing to compile).  oracle/_ref/ is listed in .gitignore (it never
enters the history) but NOT in .gpurunignore (it travels to the GPU box like the built .so files).  The run-time shims
the reference needs on a machine without librosa / a GPU live in oracle/ref_harness.py, not in the copied files.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MVB_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main() -> int:
    src_pkg = os.path.join(SRC, "fam")
    if not os.path.isdir(src_pkg):
        print(f"build_ref: {src_pkg} not found; keeping whatever oracle/_ref already holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n, h = 0, hashlib.sha256()
    for root, _dirs, files in os.walk(src_pkg):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, os.path.relpath(s, SRC))
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            h.update(open(s, "rb").read())
            n += 1
    commit = "unknown"
    sub = os.path.join(SRC, ".SUBMODULES.json")
    if os.path.isfile(sub):
        import json
        try:
            commit = json.load(open(sub)).get("commit", commit)
        except Exception:
            pass
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"metavoiceio/metavoice-src @ {commit}\n{n} files copied unmodified from {src_pkg}\nsha256(all files) = {h.hexdigest()}\n")
    print(f"build_ref: {n} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
