"""TEST INFRASTRUCTURE.  Places the UNMODIFIED reference next to the oracle so that it can be
timed (bench.py --impl reference, cpu_baseline) and compared with on a box that has no
/root/reference:

    python oracle/make_ref.py      ->  oracle/_ref/pywfo/{__init__,main,dets,helpers}.py

The four files are byte copies of /root/reference/pywfo/*.py (pure Python: nothing to compile).
oracle/_ref/ is git-ignored -- reference sources never enter this repository's history -- but it
is NOT gpurun-ignored, so the copy travels to the GPU box like a built .so.
__graft_entry__.build() runs this when /root/reference exists."""
import hashlib
import os
import shutil
import sys

REF = "/root/reference/pywfo"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "pywfo")
FILES = ("__init__.py", "main.py", "dets.py", "helpers.py")


def make_ref(verbose=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
        if verbose:
            print(f, hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()[:16])
    return True


def load_ref():
    """Import the copied reference (stdout silenced, its log file kept out of the repo).
    Returns the ``pywfo.main`` module or None when oracle/_ref is absent."""
    import contextlib
    import io
    import logging
    import tempfile

    if not os.path.exists(os.path.join(DST, "main.py")):
        return None
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())   # pywfo/__init__.py opens ./pywfo.log
    sys.path.insert(0, os.path.dirname(DST))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import pywfo.main as ref_main
        logging.getLogger("pywfo").handlers.clear()
    finally:
        sys.path.pop(0)
        os.chdir(cwd)
    return ref_main


if __name__ == "__main__":
    print("oracle/_ref written" if make_ref(verbose=True) else "no /root/reference here: nothing done")
