"""Recipe that makes the UNMODIFIED reference travel to the GPU box: packs the Python packages the caption
decode path imports (models/, misc/, config/ of yangbang18/CARE) from /root/reference into ONE archive,
oracle/_ref/reference_py.zip, which Python imports directly (zipimport).

TEST / BENCH INFRASTRUCTURE.  oracle/_ref/ is listed in .gitignore (the reference's sources never enter this
repository's tree or history) but not in .gpurunignore, so the archive rides along with the snapshot like the
built .so files.  There `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own Translator
on the host cores (`cpu_baseline.kind = "reference"`), and tests cross-check the oracle restatement against it.
Nothing is compiled: the reference is Python-only.  Run by `__graft_entry__.build()`
whenever /root/reference is present; a no-op elsewhere.

    python -m oracle.build_ref
"""
import os
import sys
import zipfile

SRC = os.environ.get("CARE_REFERENCE_SRC", "/root/reference")
DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
ARCHIVE = os.path.join(DST_DIR, "reference_py.zip")
PACKAGES = ("models", "misc", "config")


def _members():
    out = []
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = sorted(d for d in dirs if d != "__pycache__")
            for fn in sorted(files):
                if fn.endswith(".py"):
                    out.append(os.path.relpath(os.path.join(root, fn), SRC))
    return out


def build_ref(verbose=False):
    """Returns the path of the archive, or None when neither the reference nor an earlier archive is here."""
    if not os.path.isfile(os.path.join(SRC, "models", "Translator.py")):
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    members = _members()
    newest = max(os.path.getmtime(os.path.join(SRC, m)) for m in members)
    if os.path.isfile(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        with zipfile.ZipFile(ARCHIVE) as z:
            if sorted(n for n in z.namelist() if n.endswith(".py")) == sorted(members):
                return ARCHIVE
    os.makedirs(DST_DIR, exist_ok=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        # explicit directory entries: packages without an __init__.py (config/) are namespace packages, which
        # zipimport only recognises through a directory entry
        for d in sorted({os.path.dirname(m) for m in members}):
            z.writestr(d + "/", "")
        for m in members:
            z.write(os.path.join(SRC, m), m)
            if verbose:
                print("packed", m)
        z.writestr("PROVENANCE.txt", "Verbatim copy of %s/{%s}/**/*.py made by oracle/build_ref.py; not part of the "
                                     "repository history.\n" % (SRC, ",".join(PACKAGES)))
    return ARCHIVE


if __name__ == "__main__":
    print(build_ref(verbose="-v" in sys.argv))
