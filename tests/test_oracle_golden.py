"""The CPU oracle against the committed golden vectors (outputs of the imported reference, written by
oracle/make_golden.py in the build container).  Runs without a GPU and without /root/reference."""
import numpy as np
import torch

from oracle import e4s_oracle as O
from conftest import assert_close

TOL = 2e-5


def test_upfirdn2d_cases(golden):
    tags = sorted({k.split("/")[1] for k in golden if k.startswith("upfirdn2d/") and k.endswith("/cfg")})
    assert {"blur_up", "skip_up", "blur_dn", "down2", "ragged", "crop"} <= set(tags)
    for tag in tags:
        up, down, p0, p1, gain = golden[f"upfirdn2d/{tag}/cfg"]
        fir = O.make_fir((1, 3, 3, 1), float(round(gain)))
        x = torch.from_numpy(golden[f"upfirdn2d/{tag}/x"])
        y = O.upfirdn2d(x, fir, int(up), int(down), (int(p0), int(p1)))
        assert_close(y, golden[f"upfirdn2d/{tag}/y"], TOL, tag)
    x = torch.from_numpy(golden["upfirdn2d/asym/x"])
    y = O.upfirdn2d(x, torch.from_numpy(golden["upfirdn2d/asym/fir"]), 2, 1, (2, 1))
    assert_close(y, golden["upfirdn2d/asym/y"], TOL, "asym")


def test_fused_leaky_relu(golden):
    x, b = torch.from_numpy(golden["flrelu/x"]), torch.from_numpy(golden["flrelu/b"])
    y = O.fused_leaky_relu(x, b)
    assert_close(y, golden["flrelu/y"], TOL)
    gx, gb = O.fused_leaky_relu_backward(torch.from_numpy(golden["flrelu/go"]), y)
    assert_close(gx, golden["flrelu/gx"], TOL)
    assert_close(gb, golden["flrelu/gb"], TOL)


def test_generator_small(golden):
    for tag, size, K, B, nc, msz, kind in [("g64_k5", 64, 5, 2, 5, 32, "blobs"), ("g32_k13_iid", 32, 13, 1, 12, 64, "iid")]:
        st = O.synthetic_state(O.generator_param_shapes(size), salt=size)
        codes, mask, _, noise = O.synthetic_inputs(B, nc, size, msz, seed=size + K, kind=kind)
        img, feats = O.generator_forward(st, codes, mask, noise, size, K)
        assert_close(img, golden[f"generator/{tag}/image"], TOL, tag)
        assert_close(feats[:, ::16, ::2, ::2], golden[f"generator/{tag}/feats_sub"], TOL, tag)


def test_generator_grad(golden):
    size, K = 32, 13
    st = O.synthetic_state(O.generator_param_shapes(size), salt=size)
    codes, mask, _, noise = O.synthetic_inputs(1, 12, size, 64, seed=size + K, kind="iid")
    codes.requires_grad_(True)
    img, _ = O.generator_forward(st, codes, mask, noise, size, K)
    R = torch.randn(img.shape, generator=torch.Generator().manual_seed(99))
    (img * R).sum().backward()
    assert_close(codes.grad, golden["generator/g32_k13_iid/dcodes"], TOL)


def test_blocks_style_codes_and_region_mean(golden):
    for tag, (cin, cout, k, demod, up, hw) in O.MODCONV_CASES.items():
        sh = {"weight": (1, cout, cin, k, k), "modulation.weight": (cin, 512), "modulation.bias": (cin,)}
        if up:
            sh["blur.kernel"] = (4, 4)
        st = O.synthetic_state(sh, salt=len(tag))
        x, w = O.modconv_case(tag)
        y = O.modulated_conv2d(x, w, st["weight"], st["modulation.weight"], st["modulation.bias"], demod, up)
        assert_close(y, golden[f"modconv/{tag}/y"], TOL, tag)
    shapes = {}
    shapes.update(O.generator_param_shapes(64, prefix="G."))
    shapes.update(O.mlp_param_shapes(12))
    shapes.update(O.encoder_param_shapes())
    st = O.synthetic_state(shapes, salt=5)      # per-tensor seeding: G.style.* (absent here) does not matter
    sv, lat, _img, _mask = O.net3_case()
    out = O.cal_style_codes(st, sv, lat, 13)
    assert_close(out[:, :, :, ::8], golden["net3/style_codes_sub"], TOL)
    feats, m5 = O.region_mean_case()
    assert_close(O.region_mean(feats, m5), golden["region_mean/y"], TOL)


def test_mask_conversion_table(golden):
    """The 19->12 CelebAMask-HQ conversion (dataset.py:153-209) is a pure LUT."""
    from e4s_b200.masks import CELEBA19_TO_12
    for who in ("source", "target"):
        raw, c12 = golden[f"mask/{who}_raw19"], golden[f"mask/{who}_cls12"]
        assert np.array_equal(np.asarray(CELEBA19_TO_12, dtype=np.uint8)[raw], c12)


def test_gpu_baseline_structure_equals_oracle():
    """oracle/gpu_baseline.py (the reference's own execution structure: per-region loop, per-sample weights, ONE grouped
    convolution with groups = batch, model.py:287-318 - the cuDNN baseline bench.py times on the GPU) computes what the
    pinned oracle computes."""
    from oracle import gpu_baseline as GB
    size, K = 64, 5
    st = O.synthetic_state(O.generator_param_shapes(size), salt=size)
    codes, mask, _, noise = O.synthetic_inputs(2, 5, size, 32, seed=3)
    with torch.no_grad():
        a, fa = O.generator_forward(st, codes, mask, noise, size, K)
        b, fb = GB.generator_forward(st, codes, mask, noise, size, K)
    assert float((a - b).abs().max() / a.abs().max()) < 1e-5
    assert float((fa - fb).abs().max() / fa.abs().max()) < 1e-5


def test_discriminator_oracle_matches_reference_vectors():
    """oracle/disc_oracle.py against the reference Discriminator's logits (oracle/make_golden_disc.py)."""
    import os
    from conftest import ROOT
    from oracle import disc_oracle as DO
    from e4s_b200.stylegan2.model import Discriminator
    gold = np.load(os.path.join(ROOT, "tests", "golden", "disc_vectors.npz"))
    for size, batch in ((32, 4), (64, 8)):
        D = Discriminator(size).eval()                       # the mirror supplies the key layout and the registered FIR buffers
        D.load_state_dict(O.synthetic_state({k: tuple(v.shape) for k, v in D.named_parameters()}, salt=size + 1), strict=False)
        st = {k: v.detach() for k, v in D.state_dict().items()}
        with torch.no_grad():
            out = DO.discriminator_forward(st, DO.synthetic_inputs(batch, size, seed=size), size)
        assert_close(out, gold[f"d{size}/logits"], 2e-5, f"discriminator oracle {size}")
