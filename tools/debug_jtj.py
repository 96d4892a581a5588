import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyopt_b200 as tb
ctx = tb.Context(0)
np.set_printoptions(linewidth=250, precision=2, suppress=False)
rng = np.random.default_rng(0)
for (B, m, n) in [(1, 8, 256), (1, 64, 256), (1, 1000, 256), (2, 37, 200), (1, 16, 512), (1, 2048, 512)]:
    J = rng.integers(-3, 4, (B, m, n)).astype(np.float32)
    H = ctx.jtj(torch.from_numpy(J).cuda()); ctx.sync(); H = H.cpu().numpy()
    ref = np.einsum("bmi,bmj->bij", J.astype(np.float64), J.astype(np.float64))
    err = np.abs(H - ref)
    print(f"B={B} m={m} n={n}: max err {err.max():.3g} (max ref {np.abs(ref).max():.3g})")
    nb = (n + 31) // 32
    for b in range(B):
        blk = np.zeros((nb, nb))
        for i in range(nb):
            for j in range(nb):
                blk[i, j] = err[b, 32*i:32*i+32, 32*j:32*j+32].max()
        print((blk > 0).astype(int))
