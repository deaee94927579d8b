"""ORACLE support — test infrastructure only.  Builds oracle/_ref/: the UNMODIFIED reference hot path, byte-compiled.

The reference (ltdoanh2004/MotionDiffusion-MoE) is pure Python with no build system, so its
"build" is CPython byte-compilation: every module of /root/reference/text2motion/models/ is compiled from where it lies
to sourceless byte code `oracle/_ref/models/<name>.refbc` (the .pyc format under an extension that snapshot / ignore
rules for `*.pyc` do not touch; oracle/ref_runner.py imports it through importlib's SourcelessFileLoader).  No reference source is copied into the repository; oracle/_ref/ is
git-ignored (it travels to the GPU box like the repo's own built .so files).  Consumers: bench.py's `cpu_baseline` /
`--impl reference` / `gpu_eager_baseline` legs and tests (through oracle/ref_runner.py), as the thing that is timed
beside the product or the checker - never on the product path.

    python -m oracle.build_ref            # needs /root/reference; idempotent
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/text2motion/models"
DST = os.path.join(HERE, "_ref", "models")


def available():
    return os.path.exists(os.path.join(DST, "transformer.refbc"))


def build(verbose=False):
    """Returns True if oracle/_ref is present afterwards (built now, or earlier and the reference is gone)."""
    if not os.path.isdir(SRC):
        return available()
    os.makedirs(DST, exist_ok=True)
    for f in sorted(os.listdir(SRC)):
        if f.endswith(".py"):
            out = os.path.join(DST, f[:-3] + ".refbc")
            py_compile.compile(os.path.join(SRC, f), cfile=out, dfile="reference/text2motion/models/" + f, doraise=True)
            if verbose:
                print("compiled", f, "->", os.path.relpath(out, os.path.dirname(HERE)))
    with open(os.path.join(os.path.dirname(DST), "PYTHON_MAGIC"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT built (no /root/reference)")
