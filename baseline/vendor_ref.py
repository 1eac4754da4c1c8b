"""Put the UNMODIFIED reference where the GPU box can import it: ``baseline/_ref/src/...``.

``baseline/_ref/`` is git-ignored (it is not product source and never enters the history) but it is NOT
gpurun-ignored, so it travels with the snapshot like the built ``.so`` files do.  The contract's own recipe,
``pip install --target baseline/_ref /root/reference``, fails here — the reference has neither ``setup.py`` nor
``pyproject.toml`` ("Directory '/root/reference' is not installable") — so the import closure of the hot path
(``src/model``, ``src/utils``, ``src/training`` and their ``__init__.py``) is copied byte for
byte instead.  Called by ``__graft_entry__.build()`` when ``/root/reference`` exists; ``bench.py``'s reference
legs import from the copy.  Nothing under ``custom-yolo-implmentation_b200/`` ever imports it.

    python baseline/vendor_ref.py
"""
import filecmp
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("src/__init__.py", "src/model", "src/utils", "src/training")


def vendor(verbose: bool = False) -> bool:
    """Returns True when baseline/_ref holds a verified byte-identical copy."""
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.isfile(os.path.join(DST, "src", "model", "losses.py"))
    copied = []
    for rel in PACKAGES:
        src = os.path.join(REF, rel)
        files = [src] if os.path.isfile(src) else [os.path.join(d, f) for d, _, fs in os.walk(src) for f in fs if f.endswith(".py")]
        for f in files:
            out = os.path.join(DST, os.path.relpath(f, REF))
            os.makedirs(os.path.dirname(out), exist_ok=True)
            if not (os.path.exists(out) and filecmp.cmp(f, out, shallow=False)):
                shutil.copyfile(f, out)
            assert filecmp.cmp(f, out, shallow=False), out
            copied.append(out)
    if verbose:
        print(f"baseline/_ref: {len(copied)} files, byte-identical to {REF}")
    return True


if __name__ == "__main__":
    if not vendor(verbose=True):        # no /root/reference and no earlier copy: bench.py falls back to the oracle port
        print("baseline/_ref: absent (no /root/reference here); bench.py's reference legs will use the oracle port")
    sys.exit(0)
