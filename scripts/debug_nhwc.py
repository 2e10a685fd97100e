import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curvature_b200 import _native as nat
DEV = "cuda:0"
torch.manual_seed(0)
def check(name, g, prec):
    M = g.shape[1]
    X = g.reshape(g.shape[0], M, -1).permute(1, 0, 2).reshape(M, -1).double()
    want = X @ X.t()
    gd = g.to(DEV)
    if gd.dim() == 4:
        gd = gd.contiguous(memory_format=torch.channels_last)
    out = torch.zeros(M, M, device=DEV)
    nat.syrk_rows_accum(gd, False, 1.0, out, prec)
    torch.cuda.synchronize()
    o = out.cpu().double()
    print(name, prec, "equal", torch.equal(o, want), "nnz", int((o != 0).sum()), "of", o.numel(), "maxdiff", (o - want).abs().max().item())
    wsb = list(nat._workspaces.values())[0]
    part = wsb[:256 * 256 * 4].view(torch.float32).view(256, 256).cpu()
    print("   partial tile nnz", int((part[:M, :M] != 0).sum()), "part[0,:4]", part[0, :4].tolist())
    if not torch.equal(o, want):
        print(" want[0,:8]", want[0, :8].tolist()); print(" got [0,:8]", o[0, :8].tolist())
        print(" got diag[:8]", o.diag()[:8].tolist(), "want diag", want.diag()[:8].tolist())
for prec in (nat.PREC_TF32_TMA, nat.PREC_TF32):
    check("rows 64x32", torch.randint(-3, 4, (64, 32)).float(), prec)
    check("rows 8x32", torch.randint(-3, 4, (8, 32)).float(), prec)
    check("rows 256x128", torch.randint(-3, 4, (256, 128)).float(), prec)
    check("rows 100x300", torch.randint(-3, 4, (100, 300)).float(), prec)
