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
    if name.startswith("sweep"):
        # The reference's OWN sensitivity: the same problem with every input entry perturbed by 1e-15 relative
        # (seeded).  On the ill-conditioned sweep cases its gradients move by 1e-4 relative under that
        # perturbation; a parity tolerance tighter than the reference's reproducibility would test noise.
        # Stored per QP as absolute l2 differences (tests/parity.py: check_sweep allows 3x these on top of the
        # stated tolerances).
        rs = np.random.RandomState(12345)
        pp = dict(prob)
        for k in ("Q", "p", "G", "h", "A", "b"):
            v = np.asarray(prob[k], dtype=np.float64)
            pp[k] = v * (1.0 + 1e-15 * rs.randn(*v.shape)) if v.size else v
        pp["Q"] = 0.5 * (pp["Q"] + np.swapaxes(pp["Q"], -1, -2))
        r2 = ref_runner.run_reference(pp, **opts)
        B = r["zhat"].shape[0]
        out["sens_zhat"] = np.linalg.norm((r2["zhat"] - r["zhat"]).reshape(B, -1), axis=1)
        for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
            if r.get(k) is not None:
                out["sens_" + k] = np.linalg.norm((r2[k] - r[k]).reshape(B, -1), axis=1)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    """`python -m oracle.gen_golden` regenerates everything; `... --missing` only the fixtures that do not exist
    yet; `... name [name ...]` the named cases."""
    os.makedirs(OUT, exist_ok=True)
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    missing_only = "--missing" in sys.argv
    for name, (build, full_mats) in CASES.items():
        if args and name not in args:
            continue
        if missing_only and os.path.exists(os.path.join(OUT, name + ".npz")):
            continue
        pr = build()
        if "truez" in pr:                                 # test.py:88-89: dl = zhat - truez
            z = ref_runner.run_reference(dict(pr, dl=None))["zhat"]
            pr["dl"] = z - pr["truez"]
            np.save(os.path.join(OUT, name + "_dl.npy"), pr["dl"])
        save(name, pr, full_mats=full_mats)


if __name__ == "__main__":
    main()
