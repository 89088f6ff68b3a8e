"""Generate tests/golden/*.npz by running the REAL reference on seeded problems.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden
Inputs are rebuilt from `qpth_b200.problems` / the recipes below (seeded), so the
fixtures hold only an input checksum plus the reference's outputs.  For the big
configurations dQ/dG are stored as projections (dQ @ v, dG @ v with
v_k = cos(k+1)) to keep the files small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_runner                                        # noqa: E402
from oracle.cases import CASES, checksum, proj                       # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, prob, full_mats=True, **opts):
    r = ref_runner.run_reference(prob, **opts)
    d = ref_runner.run_reference_duals(prob, **opts)
    out = dict(input_checksum=checksum(prob), zhat=r["zhat"], lam=d["lam"], slacks=d["slacks"])
    if d["nus"] is not None:
        out["nus"] = d["nus"]
    for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
        g = r.get(k)
        if g is None:
            continue
        if not full_mats and g.ndim == 3:
            out[k + "_proj"] = g @ proj(g.shape[-1])
        else:
            out[k] = g
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (build, full_mats) in CASES.items():
        pr = build()
        if "truez" in pr:                                 # test.py:88-89: dl = zhat - truez
            z = ref_runner.run_reference(dict(pr, dl=None))["zhat"]
            pr["dl"] = z - pr["truez"]
            np.save(os.path.join(OUT, name + "_dl.npy"), pr["dl"])
        save(name, pr, full_mats=full_mats)


if __name__ == "__main__":
    main()
