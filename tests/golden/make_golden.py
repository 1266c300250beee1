"""Generates the committed golden fixtures with the UNMODIFIED reference tools (oracle/_ref, compiled by
oracle/Makefile from /root/reference).  Run in the build container:  python tests/golden/make_golden.py

small_k<k>_d<d>.npz holds, for a seeded synthetic text:
  image_100/101/200/201 : the four index FILES the reference tools wrote, as uint32 words
  reads, length         : ASCII reads (exact substrings, mutated reads, reads over the text ends)
  expected_std          : (L,R) printed by the reference fmIndexSearchCPU_<d>bases_<k>step      on the tag-100 file
  expected_ac           : (L,R) printed by the reference fmIndexSearchCPU_<d>bases_<k>step-ac   on the tag-200 file
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402


def reference_cli_search(workdir, k, d, ac, index_path, reads, length):
    fa = os.path.join(workdir, "reads.fa")
    helpers.write_fasta_reads(fa, reads, length)
    exe = os.path.join(helpers.REF_DIR, f"fmIndexSearchCPU_{d}bases_{k}step" + ("-ac" if ac else ""))
    helpers.run([exe, index_path, fa, str(length), str(reads.size // length)], cwd=workdir)
    out = np.loadtxt(index_path + ".res.cpu", skiprows=1, dtype=np.uint32).reshape(-1)
    os.remove(index_path + ".res.cpu")
    return out


def make_reads(text, length, num, seed):
    rng = np.random.default_rng(seed)
    exact = helpers.synth_reads(text, seed, num, length)
    mutated = exact.copy().reshape(-1, length)[: num // 4]
    pos = rng.integers(0, length, mutated.shape[0])
    mutated[np.arange(mutated.shape[0]), pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, mutated.shape[0])]
    head = text[:length]                       # prefix of the text: walks through the '$' rows
    tail = text[-length:]
    rnd = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (num // 8) * length)]
    lower = np.char.lower(exact[: 16 * length].view("S1")).view(np.uint8)   # case-insensitive decoding
    return np.concatenate([exact, mutated.reshape(-1), head, tail, rnd, lower])


def main():
    for k, d, n, length in ((1, 64, 20011, 31), (2, 64, 20011, 32), (2, 32, 6000, 20), (2, 128, 30001, 64)):
        text = helpers.synth_text(n, seed=7 + k)
        with tempfile.TemporaryDirectory() as wd:
            paths = helpers.build_reference_indexes(wd, text, k, d)
            reads = make_reads(text, length, 1536, seed=11)
            out = {"reads": reads, "length": np.int32(length), "n": np.int64(n), "k": np.int32(k), "d": np.int32(d)}
            for tag, p in paths.items():
                out[f"image_{tag}"] = np.fromfile(p, dtype=np.uint32)
            out["expected_std"] = reference_cli_search(wd, k, d, False, paths[100], reads, length)
            out["expected_ac"] = reference_cli_search(wd, k, d, True, paths[200], reads, length)
            np.savez_compressed(os.path.join(HERE, f"small_k{k}_d{d}.npz"), **out)
            same = np.array_equal(out["expected_std"], out["expected_ac"])
            print(f"k={k} d={d} n={n}: {reads.size // length} reads, std==ac: {same}")


def quirk():
    """Tiny texts whose '$' row lies in the last chunk: the reference AltCounters searcher then differs from
    the standard one.  Both outputs are recorded; the GPU path must match EACH."""
    for k, n in ((1, 124), (2, 124), (2, 100), (1, 250)):
        text = helpers.synth_text(n, seed=100 + n)
        length = 8
        reads = np.concatenate([text[i:i + length] for i in range(0, n - length + 1)])
        with tempfile.TemporaryDirectory() as wd:
            paths = helpers.build_reference_indexes(wd, text, k, 64)
            out = {"reads": reads, "length": np.int32(length), "n": np.int64(n), "k": np.int32(k), "d": np.int32(64)}
            for tag, p in paths.items():
                out[f"image_{tag}"] = np.fromfile(p, dtype=np.uint32)
            out["expected_std"] = reference_cli_search(wd, k, 64, False, paths[100], reads, length)
            out["expected_ac"] = reference_cli_search(wd, k, 64, True, paths[200], reads, length)
            assert not np.array_equal(out["expected_std"], out["expected_ac"]), "not a quirk case"
            np.savez_compressed(os.path.join(HERE, f"quirk_k{k}_n{n}.npz"), **out)
            print(f"quirk k={k} n={n}: {int((out['expected_std'] != out['expected_ac']).sum())} differing values")


if __name__ == "__main__":
    main()
    quirk()
