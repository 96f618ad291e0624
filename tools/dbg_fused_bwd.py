import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_fused import _stack_inputs, _stack_ref, LAYER_KEYS
from util_gpu import rel
from soccerdiffusion_b200.functional import EncoderStackFn, RunCfg
from soccerdiffusion_b200 import ops as O, runtime

B, S, H, L = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (5, 100, 4, 1)
x, emb_w, emb_b, pe, layers, gout = _stack_inputs(B, S, L, seed=B + S)
params = [emb_w, emb_b, *layers]
want = _stack_ref(x, emb_w, emb_b, pe, layers, B, S, H)
want.backward(gout)
ref = [p.grad.clone() for p in params]
names = ["emb_w", "emb_b"] + [f"l{l}.{k}" for l in range(L) for k in LAYER_KEYS]
for fused in (True, False):
    for p in params: p.grad = None
    runtime.set_fused_layers(fused)
    got = EncoderStackFn.apply(RunCfg(precision=O.PREC_BF16, p=0.0, seed=1, stream_base=0), B, S, H, pe, x, emb_w, emb_b, *layers)
    got.backward(gout)
    print("fused" if fused else "unfused", "out", rel(got, want))
    for n, p, g in zip(names, params, ref):
        print(f"   {n:12s} rel {rel(p.grad, g):.4f}  |ref| {float(g.norm()):.3g}")
    if fused:
        g_ref = ref[2 + LAYER_KEYS.index("in_w")]
        g_got = params[2 + LAYER_KEYS.index("in_w")].grad
        dh = 128 // H
        for blk, nm in enumerate("qkv"):
            print("   in_w", nm, " ".join(f"h{h}:{rel(g_got[blk*128+h*dh: blk*128+(h+1)*dh], g_ref[blk*128+h*dh: blk*128+(h+1)*dh]):.3f}" for h in range(H)))
