import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import case_inputs, oracle_kwargs, npy
from relightable3dgaussian_b200 import _C_raster as C
from oracle import oracle
import test_raster_gpu as T
W, H, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sc, cam = case_inputs(P, W, H, 0, view=1, scale_boost=float(sys.argv[4]))
kw = oracle_kwargs(sc, cam, torch.tensor([0.1, 0.0, 0.3]))
extra = dict(shs=npy(sc.shs), scales=npy(sc.scales), rotations=npy(sc.rotations))
o = T.run_ours(C, pseudo=False, **kw, **extra)
f = oracle.rasterize_forward(computer_pseudo_normal=False, **kw, **extra)
d = np.abs(npy(o["color"]) - f["img"]["color"]).max(axis=0)
bad = np.argwhere(d > 1e-4)
print("bad pixels", len(bad), "max", d.max())
nc_o = npy(o["n_contrib"]); nc_f = f["img"]["n_contrib"].reshape(H, W)
print("n_contrib mismatches", (nc_o != nc_f).sum())
for (y, x) in bad[:10]:
    print("pix", x, y, "tile", (y // 16) * ((W + 15) // 16) + x // 16, "ours", npy(o["color"])[:, y, x], "oracle", f["img"]["color"][:, y, x], "ncontrib", nc_o[y, x], nc_f[y, x],
          "T ours", float(npy(o["mid"]("final_T")).reshape(H, W)[y, x]))
if len(bad):
    ys, xs = bad[:, 0], bad[:, 1]
    print("bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max())
    # blocks 8x4
    blk = set((int(y) // 4, int(x) // 8) for y, x in bad)
    print("distinct 8x4 blocks", len(blk), "tiles", len(set((int(y)//16, int(x)//16) for y, x in bad)))
r = npy(o["radii"]); print("radii max", r.max(), "mean", r[r > 0].mean())
