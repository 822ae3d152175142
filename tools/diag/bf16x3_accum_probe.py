"""Splits the error of the bf16x3 tcgen05 GEMM into (a) operand representation + dropped lo*lo term and (b) the
tensor-core fp32 accumulation itself, as a function of the accumulation length K."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rel_pose_b200 import ops

def f64(planes):
    return planes.float().cpu().numpy().astype(np.float64)

torch.manual_seed(0)
M, N = 256, 192
for positive in (True, False):
    for K in (192, 768, 3072, 12288):
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / np.sqrt(K)
        if positive:
            A = A.abs()                      # post-ReLU activations
        exact = A.double().cpu().numpy() @ W.double().cpu().numpy().T
        Ap, Wp = ops.split_planes(A, 2), ops.split_planes(W, 2)
        got = ops.linear_tc(Ap, Wp)[0].cpu().numpy().astype(np.float64)
        a, w = f64(Ap), f64(Wp)
        emu = a[0] @ w[0].T + a[0] @ w[1].T + a[1] @ w[0].T          # the three products, exact accumulation
        simt = ops.linear(A, W).cpu().numpy().astype(np.float64)
        rms = np.sqrt((exact ** 2).mean())
        rel = lambda e: np.sqrt((e ** 2).mean()) / rms
        shrink = ((got - emu) * np.sign(emu)).mean() / rms
        print(f"positive_A={positive!s:5s} K={K:6d}: total {rel(got - exact):.2e}  representation {rel(emu - exact):.2e}  "
              f"accumulation {rel(got - emu):.2e} (signed toward zero: {shrink:+.2e})  fp32 SIMT {rel(simt - exact):.2e}")
