"""Child process of tests/test_gpu_zz_fallback_families.py: runs every job of tests/fallback_jobs.py through QPFunction on
cuda:0 and writes <out_dir>/<job>.npz (or <job>.err with the traceback). Usage: python -m tests.gpu_child <out_dir>."""
import os
import sys
import traceback

import numpy as np


def _run_eq_only(cfg, dev="cuda:0"):
    import torch
    from qpth_b200 import QPFunction
    from tests.fallback_jobs import eq_only_problem
    pr = eq_only_problem(**cfg)
    t = {k: torch.tensor(pr[k], dtype=torch.float64, device=dev, requires_grad=True) for k in ("Q", "p", "A", "b")}
    e = torch.Tensor().to(dev).double()
    z = QPFunction(verbose=-1)(t["Q"], t["p"], e, e, t["A"], t["b"])
    z.backward(torch.tensor(pr["dl"], dtype=torch.float64, device=dev))
    return dict(zhat=z.detach().cpu().numpy(), grads=[t[k].grad.cpu().numpy() for k in ("Q", "p", "A", "b")])


def _save(out_dir, name, out):
    rec = {k: np.asarray(out[k]) for k in ("zhat", "lam", "slacks", "iters") if out.get(k) is not None}
    if out.get("nus") is not None:
        rec["nus"] = np.asarray(out["nus"])
    for i, g in enumerate(out.get("grads") or ()):
        if g is not None:
            rec["grad%d" % i] = np.asarray(g)
    np.savez(os.path.join(out_dir, name + ".npz"), **rec)


def main(out_dir):
    from oracle.cases import load_case
    from qpth_b200 import qp as qpmod
    from qpth_b200.problems import random_qp_batch
    from tests.fallback_jobs import jobs
    from tests.test_gpu_parity import _run
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, kind, payload, env, mode in jobs():
        saved = {k: os.environ.get(k) for k in env}
        try:
            os.environ.update(env)
            qpmod.MODE = mode or "auto"
            if kind == "eq_only":
                out = _run_eq_only(payload)
            else:
                prob = load_case(payload, golden)[0] if kind == "golden" else random_qp_batch(**payload)
                out = _run(prob)
            _save(out_dir, name, out)
        except BaseException:      # noqa: BLE001 - recorded for the parent, the next job still runs
            with open(os.path.join(out_dir, name + ".err"), "w") as fh:
                fh.write(traceback.format_exc())
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        with open(os.path.join(out_dir, "progress.txt"), "a") as fh:
            fh.write(name + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
