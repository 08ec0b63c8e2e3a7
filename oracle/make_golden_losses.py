#!/usr/bin/env python
"""Pins oracle/loss_oracle.py to the REAL reference loss code and writes tests/golden/loss_vectors.npz.

Run in the build container (needs /root/reference; never on the GPU box):

    python oracle/make_golden_losses.py

Imports the reference's own classes - src.criteria.lpips.lpips.LPIPS, src.criteria.id_loss.IDLoss,
src.criteria.face_parsing.face_parsing_loss.FaceParsingLoss - and executes the unmodified source text of
Optimizer.calc_loss (scripts/optimization.py:88-122) on them.  The only interventions are the ones the missing network
forces: torchvision's pretrained AlexNet download and the LPIPS linear-layer download (networks.py:77, utils.py:11-19) are
replaced by seeded state dicts, IDLoss reads a seeded IR-SE50 state dict from a temporary file (id_loss.py:15), and
torch.load maps the shipped parsing checkpoint (saved from CUDA tensors) to the CPU because this container has no GPU.  The
parsing UNet is pinned twice: with seeded weights, and with the checkpoint the reference ships
(pretrained_ckpts/auxiliray/model.pth, used as is).  For every case: assert oracle == reference, store the REFERENCE's
outputs.  Stored: per-network losses, calc_loss terms at the reference's own scales (1024 / 512 / 256), the gradient of
calc_loss with respect to the reconstruction (what the generator's backward receives).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import loss_oracle as LO  # noqa: E402

TOL = 2e-5


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_reference(states):
    import torchvision
    from src.criteria.lpips import lpips as ref_lpips_mod, networks as ref_networks
    # no network: an untrained torchvision AlexNet of the same architecture, and seeded linear layers
    real_alexnet = torchvision.models.alexnet
    ref_networks.models = types.SimpleNamespace(alexnet=lambda pretrained=False: real_alexnet(weights=None))
    ref_lpips_mod.get_state_dict = lambda net_type="alex", version="0.1": {k[len("lin."):]: v for k, v in states["lpips"].items() if k.startswith("lin.")}
    lp = ref_lpips_mod.LPIPS(net_type="alex").eval()
    lp.load_state_dict(states["lpips"], strict=True)

    from src.criteria.id_loss import IDLoss
    tmp = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
    torch.save({k[len("facenet."):]: v for k, v in states["id"].items()}, tmp.name)
    opts = types.SimpleNamespace(ir_se50_path=tmp.name, id_loss_multiscale=True,
                                 face_parsing_model_path=os.path.join(REF, "pretrained_ckpts", "auxiliray", "model.pth"))
    idl = IDLoss(opts).eval()
    os.unlink(tmp.name)

    from src.criteria.face_parsing.face_parsing_loss import FaceParsingLoss
    real_load = torch.load                    # the shipped checkpoint holds CUDA tensors and this container has no GPU
    torch.load = lambda f, *a, **k: real_load(f, *a, **{**k, "map_location": "cpu"})
    try:
        fp_real = FaceParsingLoss(opts).eval()                                # the shipped checkpoint, as is
        fp_seed = FaceParsingLoss(opts).eval()
    finally:
        torch.load = real_load
    fp_seed.load_state_dict(states["parsing"], strict=True)
    return lp, idl, fp_seed, fp_real, opts


def reference_calc_loss():
    """The unmodified source text of Optimizer.calc_loss, bound to a stub object."""
    src = open(os.path.join(REF, "scripts", "optimization.py")).read()
    a = src.index("    def calc_loss(self, img, img_recon, mask):")
    b = src.index("    def setup_W_optimizer", a)
    body = "\n".join(line[4:] if line.startswith("    ") else line for line in src[a:b].splitlines())
    ns = {"F": F, "torch": torch}
    exec(body, ns)
    return ns["calc_loss"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    states = LO.loss_states(salt=11)
    lp, idl, fp_seed, fp_real, opts = build_reference(states)
    real_parsing = {"G." + k: v for k, v in torch.load(opts.face_parsing_model_path, map_location="cpu").items()}
    out, worst = {}, 0.0

    def check(name, ours, ref):
        nonlocal worst
        e = rel(ours, ref)
        worst = max(worst, e)
        print(f"{name:40s} oracle vs reference: {e:.2e}")
        assert e <= TOL, (name, e)
        out[name] = np.asarray(torch.as_tensor(ref).detach().float().numpy())

    img, recon, far = LO.golden_inputs()          # seeded; tests regenerate them the same way

    with torch.no_grad():
        for tag, r in (("near", recon), ("far", far)):
            check(f"lpips/{tag}", LO.lpips(states["lpips"], r, img), lp(r, img))
            check(f"id/{tag}", LO.id_loss(states["id"], r, img), idl(r, img)[0])
            check(f"parsing_seeded/{tag}", LO.parsing_loss(states["parsing"], r, img), fp_seed(r, img)[0])
            check(f"parsing_shipped/{tag}", LO.parsing_loss(real_parsing, r, img), fp_real(r, img)[0])
        feats_ref = idl.extract_feats(img)
        for i, (a, b) in enumerate(zip(LO.id_extract_feats(states["id"], img), feats_ref)):
            check(f"id/feats{i}", a[:, :4096], b[:, :4096])
        for i, (a, b) in enumerate(zip(LO.parsing_extract_feats(real_parsing, img), fp_real.extract_feats(img))):
            check(f"parsing_shipped/feats{i}", a[:, :4096], b[:, :4096])

    # calc_loss at the reference's own scales (adaptive pooling of the 256x256 pair to 1024 / 512 / 256), value and gradient
    calc = reference_calc_loss()
    stub = types.SimpleNamespace(opts=types.SimpleNamespace(id_lambda=0.1, l2_lambda=1.0, lpips_lambda=0.8, face_parsing_lambda=0.1),
                                 id_loss=idl, lpips_loss=lp, face_parsing_loss=fp_seed)
    r_ref = recon[:1].clone().requires_grad_(True)
    loss_ref, dict_ref, _ = calc(stub, img[:1], r_ref, None)
    loss_ref.backward()
    r_or = recon[:1].clone().requires_grad_(True)
    loss_or, terms = LO.calc_loss(states, img[:1], r_or)
    loss_or.backward()
    check("calc_loss/loss", loss_or.detach(), loss_ref.detach())
    for k in ("loss_id", "loss_l2", "loss_lpips", "loss_face_parsing"):
        check(f"calc_loss/{k}", terms[k].detach(), torch.tensor(dict_ref[k]))
    check("calc_loss/grad_recon_full", r_or.grad, r_ref.grad)
    out["calc_loss/grad_recon"] = out.pop("calc_loss/grad_recon_full")[:, :, ::4, ::4].copy()       # every 4th pixel is kept

    dst = os.path.join(ROOT, "tests", "golden", "loss_vectors.npz")
    np.savez_compressed(dst, **out)
    print(f"worst oracle-vs-reference error {worst:.2e}; wrote {dst} ({os.path.getsize(dst) / 1e6:.2f} MB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
