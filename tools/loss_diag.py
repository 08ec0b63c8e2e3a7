#!/usr/bin/env python
"""Diagnostic: loss networks on the GPU (e4s_b200.criteria, cuDNN) against the CPU oracle under different cuDNN settings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import loss_oracle as LO
from e4s_b200.criteria import InversionLoss
from e4s_b200.synthetic import load_synthetic_losses

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())

m = InversionLoss(); load_synthetic_losses(m, 11); m = m.cuda()
st = LO.loss_states(11)
img, recon, far = LO.golden_inputs()
print("torch", torch.__version__, "cudnn", torch.backends.cudnn.version())
print("cudnn.allow_tf32", torch.backends.cudnn.allow_tf32, "matmul.allow_tf32", torch.backends.cuda.matmul.allow_tf32)
for attr in ("fp32_precision",):
    for mod in (torch.backends, torch.backends.cudnn, getattr(torch.backends.cudnn, "conv", None), torch.backends.cuda.matmul):
        if mod is not None and hasattr(mod, attr):
            print(" ", mod, attr, getattr(mod, attr))
with torch.no_grad():
    ref_id = LO.id_extract_feats(st["id"], img)
    ref_par = LO.parsing_extract_feats(st["parsing"], img)
    ref_lp = LO.alexnet_features(st["lpips"], img)
from e4s_b200.criteria.inversion_loss import conv_precision
with torch.no_grad(), conv_precision(True):
    a = m.id_loss.extract_feats(img.cuda()); b = m.face_parsing_loss.extract_feats(img.cuda()); c = m.lpips_loss.features(img.cuda())
print(f"{'conv_precision(exact)':36s} id feats", " ".join(f"{rel(x, y):.1e}" for x, y in zip(a, ref_id)), "| parsing", " ".join(f"{rel(x, y):.1e}" for x, y in zip(b, ref_par)),
      "| lpips", " ".join(f"{rel(x, y):.1e}" for x, y in zip(c, ref_lp)))
r = recon[:1].cuda().requires_grad_(True)
with conv_precision(True):
    l = m.id_loss.loss_from_feats(m.id_loss.extract_feats(r), [f.detach() for f in m.id_loss.extract_feats(img[:1].cuda())])
    l.backward()
rc = recon[:1].clone().requires_grad_(True)
lc = LO.id_loss(st["id"], rc, img[:1]); lc.backward()
print(f"{'conv_precision(exact)':36s} id loss {float(l):.6f} vs {float(lc):.6f}; grad rel err {rel(r.grad, rc.grad):.2e}")
settings = [("default", dict()), ("allow_tf32=False", dict(allow_tf32=False)), ("allow_tf32=False,deterministic", dict(allow_tf32=False, deterministic=True)),
            ("cudnn disabled", dict(enabled=False))]
for name, kw in settings:
    with torch.no_grad(), torch.backends.cudnn.flags(**({"enabled": True} | kw)):
        a = m.id_loss.extract_feats(img.cuda())
        b = m.face_parsing_loss.extract_feats(img.cuda())
        c = m.lpips_loss.features(img.cuda())
    print(f"{name:36s} id feats", " ".join(f"{rel(x, y):.1e}" for x, y in zip(a, ref_id)), "| parsing", " ".join(f"{rel(x, y):.1e}" for x, y in zip(b, ref_par)),
          "| lpips", " ".join(f"{rel(x, y):.1e}" for x, y in zip(c, ref_lp)))
# gradient of the id loss wrt the reconstruction
for name, kw in settings:
    r = recon[:1].cuda().requires_grad_(True)
    with torch.backends.cudnn.flags(**({"enabled": True} | kw)):
        l = m.id_loss.loss_from_feats(m.id_loss.extract_feats(r), [f.detach() for f in m.id_loss.extract_feats(img[:1].cuda())])
        l.backward()
    rc = recon[:1].clone().requires_grad_(True)
    lc = LO.id_loss(st["id"], rc, img[:1]); lc.backward()
    print(f"{name:36s} id loss {float(l):.6f} vs {float(lc):.6f}; grad rel err {rel(r.grad, rc.grad):.2e}")
# the per-unit error growth inside IR-SE50 (TF32 off)
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
    x = torch.nn.functional.adaptive_avg_pool2d(img[:, :, 35:223, 32:220], (112, 112))
    xg = x.cuda()
    net = m.id_loss.facenet
    yg = net.input_layer(xg)
    import torch.nn.functional as F
    sd = st["id"]; p = "facenet."
    yc = F.prelu(LO._bn(sd, p + "input_layer.1", F.conv2d(x, sd[p + "input_layer.0.weight"], None, 1, 1)), sd[p + "input_layer.2.weight"])
    print("input layer", f"{rel(yg, yc):.1e}")
    for i, unit in enumerate(net.body):
        yg_new = unit(yg)
        # CPU on the GPU's input: isolates this unit's own error
        yc_same = unit.cpu()(yg.cpu()); unit.cuda()
        print(f"unit {i:2d} own error {rel(yg_new, yc_same):.1e}  (|y| max {float(yg_new.abs().max()):.2e})")
        yg = yg_new
