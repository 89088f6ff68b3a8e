"""Golden vectors for the fused OptNet layer: the notebook's block (example-cls-layer.ipynb:125-130) with the
REAL reference's QPFunction on CPU, forward + backward. TEST INFRASTRUCTURE ONLY; build container only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref_qp, _ = ref_runner.load()
    for name, (seed, B, n, m) in {"layer_small": (1, 6, 20, 30), "layer_cls": (2, 16, 60, 60)}.items():
        g = torch.Generator().manual_seed(seed)
        L = torch.tril(torch.rand(n, n, generator=g, dtype=torch.float64)).requires_grad_(True)
        G = (torch.rand(m, n, generator=g, dtype=torch.float64) * 2 - 1).requires_grad_(True)
        z0 = (0.1 * torch.randn(n, generator=g, dtype=torch.float64)).requires_grad_(True)
        s0 = (1.0 + torch.rand(m, generator=g, dtype=torch.float64)).requires_grad_(True)
        p = torch.randn(B, n, generator=g, dtype=torch.float64).requires_grad_(True)
        dl = torch.randn(B, n, generator=g, dtype=torch.float64)
        eps = 1e-4
        Lm = torch.tril(torch.ones(n, n, dtype=torch.float64)) * L
        Q = Lm.mm(Lm.t()) + eps * torch.eye(n, dtype=torch.float64)
        h = G.mv(z0) + s0
        e = torch.Tensor().double()
        z = ref_qp.QPFunction(verbose=-1)(Q, p, G, h, e, e)
        z.backward(dl)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), L=L.detach().numpy(), G=G.detach().numpy(), z0=z0.detach().numpy(),
                            s0=s0.detach().numpy(), p=p.detach().numpy(), dl=dl.numpy(), eps=np.array(eps), z=z.detach().numpy(),
                            dL=L.grad.numpy(), dG=G.grad.numpy(), dz0=z0.grad.numpy(), ds0=s0.grad.numpy(), dp=p.grad.numpy())
        print(name, "z", float(z.abs().max()), "dL", float(L.grad.abs().max()))


if __name__ == "__main__":
    main()
