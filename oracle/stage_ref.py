"""TEST INFRASTRUCTURE ONLY.  Stages the four reference modules the zero-edit boundary test imports — byte-for-byte, from
where they lie under $REFERENCE (default /root/reference) — into the git-ignored oracle/_ref/ref_py/ so that they travel to
the GPU box with the snapshot (the box has no /root/reference).  Nothing of the reference enters the repository history;
nothing in the product imports from here.

    python oracle/stage_ref.py        (also run by __graft_entry__.build() when the reference is present)

Staged: scripts/utils/{pq_utils,Singleton,Namespace,Timer}.py — pq_utils.py:1-6 imports the other three.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["pq_utils.py", "Singleton.py", "Namespace.py", "Timer.py"]


def stage(reference=None) -> str:
    reference = reference or os.environ.get("REFERENCE", "/root/reference")
    src = os.path.join(reference, "scripts", "utils")
    dst_root = os.path.join(HERE, "_ref", "ref_py")
    dst = os.path.join(dst_root, "scripts", "utils")
    os.makedirs(dst, exist_ok=True)
    for pkg in (os.path.join(dst_root, "scripts"), dst):
        open(os.path.join(pkg, "__init__.py"), "a").close()
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return dst_root


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else None))
