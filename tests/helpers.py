"""Shared test plumbing: loads the product library, the C oracle and the
reference tools (oracle/_ref), and builds small seeded data sets with the
UNMODIFIED reference index builders.

Nothing here reads /root/reference at run time: oracle/_ref holds binaries that
were compiled from it in the build container and travel with the repo.
"""
import ctypes as C
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
PKG_NAME = "k-step_fm-index_b200"
FMSYNTH = os.path.join(ROOT, PKG_NAME, "bin", "fmsynth")
TAG_SUFFIX = {100: "", 101: ".interleaving", 200: ".ac", 201: ".interleaving.ac"}

_built = False


def ensure_built():
    global _built
    if _built:
        return
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    _built = True


def pkg():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module(PKG_NAME)


def has_ref_tools():
    return os.path.exists(os.path.join(REF_DIR, "gfmiBaseLine_64bases_2step"))


sys.path.insert(0, ORACLE_DIR)
from bindings import Oracle, RefSearcher  # noqa: E402,F401  (oracle/bindings.py)


# --------------------------------------------------------------------------- #
# data sets
# --------------------------------------------------------------------------- #
def run(cmd, cwd=None):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, f"{cmd} failed:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
    return p.stdout


def synth_text(n, seed):
    """The text of fm_synth.h as a numpy uint8 array of ASCII bases (vectorised restatement)."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[(x >> np.uint64(62)).astype(np.int64)]


def synth_read_starts(seed, num, n, length, first=0):
    j = np.arange(first, first + num, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = ((np.uint64(seed) ^ np.uint64(0xA5A5A5A55A5A5A5A)) + j + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x % np.uint64(n - length + 1)).astype(np.int64)


def synth_reads(text, seed, num, length, first=0):
    starts = synth_read_starts(seed, num, text.size, length, first)
    return text[starts[:, None] + np.arange(length)[None, :]].reshape(-1)


def write_fasta_ref(path, text):
    with open(path, "wb") as f:
        f.write(b"> %d" % text.size)
        for i in range(0, text.size, 70):
            f.write(b"\n" + text[i:i + 70].tobytes())
        f.write(b"\n")


def write_fasta_reads(path, reads, length):
    r = np.ascontiguousarray(reads, dtype=np.uint8).reshape(-1, length)
    with open(path, "wb") as f:
        for i in range(r.shape[0]):
            f.write(b">rid%d\n" % (i + 1) + r[i].tobytes() + b"\n")


def build_reference_indexes(workdir, text, k, d, name="ref.fa"):
    """Runs the reference builder + both transformers on `text`; returns {tag: path}."""
    os.makedirs(workdir, exist_ok=True)
    fa = os.path.join(workdir, name)
    write_fasta_ref(fa, text)
    n = text.size
    run([os.path.join(REF_DIR, f"gfmiBaseLine_{d}bases_{k}step"), name, str(n)], cwd=workdir)
    base = f"{fa}.{n}.{d}fmi{k}steps.fmi"
    run([os.path.join(REF_DIR, f"tfmiBMP_{d}bases_{k}step"), os.path.basename(base)], cwd=workdir)
    run([os.path.join(REF_DIR, f"tfmiAC_{d}bases_{k}step"), os.path.basename(base)], cwd=workdir)
    return {tag: base + suf for tag, suf in TAG_SUFFIX.items()}


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def results_text_md5(res):
    """md5 of the reference's text dump (common/common.c:201-220) of an (L,R) array."""
    r = np.asarray(res, dtype=np.uint32).reshape(-1, 2)
    h = hashlib.md5()
    h.update(b"%d\n" % r.shape[0])
    for i in range(0, r.shape[0], 1 << 16):
        h.update("".join(f"{a} {b}\n" for a, b in r[i:i + (1 << 16)].tolist()).encode())
    return h.hexdigest()
