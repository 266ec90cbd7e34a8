"""Diagnostic: how far is a bf16-autocast evaluation of the SAME graph (oracle run on the GPU under autocast, LayerNorm
kept in fp32 like torch's autocast policy) from the fp32 oracle, next to our bf16 run.  Sets the bf16 tolerance."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import qavit_oracle as O
from util import build_model, inputs, rel_max, rel_l2
import qavit_b200 as Q

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
_ln32 = O._ln
def _ln_autocast(x, sd, prefix, eps=1e-5):
    return _ln32(x.float(), sd, prefix, eps)
for case in ["hqavit_c100", "qavitv2_c100"]:
    for B in (4, 32):
        model, ocfg, sd, _ = build_model(case, precision="bf16")
        model.train()
        x, y = inputs(ocfg, B)
        ref_logits, ref_loss, grads, _ = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.1)
        # emulated autocast reference on GPU
        sdg = {k: v.cuda() for k, v in sd.items()}
        O._ln = _ln_autocast
        with torch.autocast("cuda", dtype=torch.bfloat16):
            emu_logits, emu_loss, emu_grads, _ = O.loss_and_grads(sdg, ocfg, x.cuda(), y.cuda(), label_smoothing=0.1)
        O._ln = _ln32
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(x.cuda())
        loss = Q.cross_entropy(logits, y.cuda(), label_smoothing=0.1); loss.backward()
        def gerr(gd, get):
            num = den = 0.0
            for n, g in grads.items():
                if g is None: continue
                d = (get(n).float().cpu() - g).norm().item(); num += d*d; den += g.norm().item()**2
            return (num/den)**0.5
        named = dict(model.named_parameters())
        print(f"{case} B={B}: logits rel_max ours {rel_max(logits, ref_logits):.3e} emu {rel_max(emu_logits.float(), ref_logits):.3e} | "
              f"rel_l2 ours {rel_l2(logits, ref_logits):.3e} emu {rel_l2(emu_logits.float(), ref_logits):.3e} | loss ours {loss.item():.5f} emu {emu_loss.item():.5f} ref {ref_loss.item():.5f} | "
              f"grads ours {gerr(grads, lambda n: named[n].grad):.3e} emu {gerr(grads, lambda n: emu_grads[n]):.3e}")
        # blocks bf16 but lateral path fp32 (no autocast)
        if case == "hqavit_c100":
            model.zero_grad(set_to_none=True)
            model2, _, _, _ = build_model(case, precision="bf16"); model2.train()
            lg2 = model2(x.cuda())
            print(f"   blocks bf16 + lateral fp32: logits rel_max {rel_max(lg2, ref_logits):.3e}")
            model3, _, _, _ = build_model(case, precision="fp32"); model3.train()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                lg3 = model3(x.cuda())
            print(f"   blocks fp32 + lateral autocast: logits rel_max {rel_max(lg3, ref_logits):.3e}")
