"""Import the UNMODIFIED reference in the build container (thin alias of oracle/ref_loader.py, kept for the tests
and the golden generators that were written against this name)."""
from oracle.ref_loader import available, build_nets, iteration, load, ref_dir  # noqa: F401

REF_DIR = ref_dir()
