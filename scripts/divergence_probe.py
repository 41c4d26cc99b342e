"""Debug probe: run Stage-I steps at a given batch on cuda:0 and report non-finite gradients / losses per step."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from thesis_fmri_reconstruction_b200 import engine, hp, init  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P, S = init.init_vaegan(hp.CFG64, 128, seed=12345)
tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
g = torch.Generator().manual_seed(1234)
x = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).cuda()
e = torch.randn(B, 128, generator=g).cuda()
zp = torch.randn(B, 128, generator=g).cuda()
for i in range(steps):
    out = tr.forward_backward(x, e, zp)
    torch.cuda.synchronize()
    bad_f = [k for k, v in out.items() if torch.is_tensor(v) and not torch.isfinite(v.float()).all()]
    bad_g = [(k, int((~torch.isfinite(v)).sum())) for k, v in tr.named_grads().items() if not torch.isfinite(v).all()]
    gn = {b: float(torch.nan_to_num(bk.flat_g).norm()) for b, bk in tr.buckets.items()}
    tr.update(B)
    torch.cuda.synchronize()
    bad_p = [k for k, v in tr.named_parameters().items() if not torch.isfinite(v).all()]
    lo = tr.losses()
    print(i, "non-finite forward:", bad_f, "| grads:", bad_g[:6], "| params:", bad_p[:6], "| grad norms:", gn,
          "| kl/B", lo["kl"] / B, "mse/B", lo["mse"] / B)
