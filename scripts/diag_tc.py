import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from gan_testlib import rel_l2
import test_tc_vs_simt_gpu as T
T.ACT_BUFFERS_D = [("d.h1", torch.bfloat16), ("d.h2", torch.bfloat16), ("d.h3", torch.bfloat16), ("d.pool", torch.float32), ("d.dz3", torch.bfloat16),
                   ("d.dz2", torch.bfloat16), ("d.dz1", torch.bfloat16), ("d.gx", torch.float32), ("g.y0", torch.bfloat16), ("g.x1", torch.bfloat16), ("g.y1", torch.bfloat16), ("g.x2", torch.bfloat16), ("g.notes", torch.float32)]
for B in (8, 130):
    for which in ("d", "g"):
        m0, b0, g0 = T._run(B, which, False)
        m1, b1, g1 = T._run(B, which, True)
        print(f"== B={B} {which}: metrics simt {m0.tolist()} tc {m1.tolist()}")
        for n in b0:
            print(f"   buf {n:10s} {rel_l2(b1[n], b0[n]):.3e}")
        for k in g0:
            print(f"   grad {k:40s} {rel_l2(g1[k], g0[k]):.3e}")
