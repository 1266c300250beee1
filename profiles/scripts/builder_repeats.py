"""GPU builder on a large repetitive text (prefix-doubling path): time, and bytes vs the reference builder."""
import hashlib, importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
pkg = helpers.pkg()
unit = helpers.synth_text(1_000_000, 3)
mut = unit.copy(); mut[::5003] = ord("A")                      # diverged copy of the repeat
text = np.concatenate([unit, helpers.synth_text(500_000, 4), unit, mut, np.full(40_000, ord("T"), dtype=np.uint8), unit[:700_000],
                       np.tile(np.frombuffer(b"ACACGT", dtype=np.uint8), 50_000), helpers.synth_text(15_000_000, 5), mut])
n = text.size
for k in (2,):
    t0 = time.time(); b = pkg.IndexBuild.from_text(text, k, 64); img = b.download(); t1 = time.time(); b.free()
    print(json.dumps({"what": "gpu build (repetitive text)", "n": int(n), "k": k, "seconds": t1 - t0, "md5": hashlib.md5(img.data).hexdigest()}), flush=True)
    with tempfile.TemporaryDirectory() as wd:
        t0 = time.time(); paths = helpers.build_reference_indexes(wd, text, k, 64); t1 = time.time()
        want = np.fromfile(paths[100], dtype=np.uint32)
    print(json.dumps({"what": "reference gfmiBaseLine", "seconds": t1 - t0, "identical": bool(np.array_equal(img, want))}), flush=True)
