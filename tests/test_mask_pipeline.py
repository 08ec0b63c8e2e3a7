"""Mask stage of the face-swapping pipeline: shape swapping, foreground mask, blending masks,
texture-vector swap.  Integer / index work - compared BIT-EXACTLY.

CPU part: oracle/mask_oracle.py against tests/golden/mask_pipeline_vectors.npz, the outputs of the reference's own
functions (oracle/make_golden_masks.py).  GPU part: the CUDA kernels through the C ABI against the same vectors and
against the oracle on seeded random inputs, ragged sizes and the empty / constant edge cases.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import mask_oracle as MO
from oracle.make_golden_masks import synthetic_label_maps


@pytest.fixture(scope="module")
def mgold():
    d = np.load(os.path.join(ROOT, "tests", "golden", "mask_pipeline_vectors.npz"))
    return {k: d[k] for k in d.files}


def _cases(golden):
    src, tgt = golden["mask/source_cls12"].astype(np.uint8), golden["mask/target_cls12"].astype(np.uint8)
    rnd = synthetic_label_maps(11, 4, 96, 72)
    return {"example": (src, tgt), "example_rev": (tgt, src), "random_a": (rnd[0], rnd[1]), "random_b": (rnd[2], rnd[3])}


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_matches_reference_vectors(golden, mgold):
    for tag, (s, t) in _cases(golden).items():
        for hf in (True, False):
            res, hole = MO.swap_head_mask(s, t, hf)
            assert np.array_equal(res, mgold[f"swap/{tag}/hair{int(hf)}/res"]), (tag, hf)
            assert np.array_equal(hole, mgold[f"swap/{tag}/hair{int(hf)}/hole"]), (tag, hf)
        res, hole = MO.swap_head_mask(s, t)
        fg = MO.foreground_mask(res, hole)
        assert np.array_equal(fg, mgold[f"fg/{tag}"])
        for r in (0, 1, 5):
            for op in ("dilation", "erosion", "expansion"):
                _, border, full = MO.create_masks(fg, r, op)
                assert np.array_equal(border, mgold[f"masks/{tag}/r{r}/{op}/border"]), (tag, r, op)
                assert np.array_equal(full, mgold[f"masks/{tag}/r{r}/{op}/full"]), (tag, r, op)


def test_oracle_style_vector_swap(mgold):
    comp = sorted(set(range(12)) - {0, 4, 11, 10})
    for tag in ("plain", "no_ear", "no_teeth", "neither"):
        for interp in (0, 1):
            k = f"stylevec/{tag}/interp{interp}"
            assert np.array_equal(MO.swap_comp_style_vector(mgold[k + "/sv1"], mgold[k + "/sv2"], comp, bool(interp)), mgold[k + "/out"])


def test_swap_properties():
    """Size-independent properties of the shape swap: idempotent labels stay inside the 12 classes, the target's
    background survives, holes are exactly the pixels no rule claimed."""
    maps = synthetic_label_maps(3, 2, 200, 136)
    res, hole = MO.swap_head_mask(maps[0], maps[1])
    assert res.max() <= 11 and set(np.unique(hole)) <= {0, 255}
    assert np.array_equal(res == 0, maps[1] == 0)                       # background of the target, nothing else
    assert np.all(res[hole == 255] == 6)
    same, hole_same = MO.swap_head_mask(maps[1], maps[1])                # swapping a face with itself
    keep = np.isin(maps[1], (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11))
    assert np.array_equal(same[keep & (hole_same == 0)], maps[1][keep & (hole_same == 0)])


# ------------------------------------------------------------------------------------------------ GPU: kernels
def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
def test_swap_head_mask_kernel_golden(golden, mgold):
    from e4s_b200.masks import swap_head_mask_revisit_considerGlass, swap_head_mask_with_foreground
    for tag, (s, t) in _cases(golden).items():
        for hf in (True, False):
            res, hole = swap_head_mask_revisit_considerGlass(cu(s), cu(t), hair_first=hf)
            assert np.array_equal(res.cpu().numpy(), mgold[f"swap/{tag}/hair{int(hf)}/res"]), (tag, hf)
            assert np.array_equal(hole.cpu().numpy(), mgold[f"swap/{tag}/hair{int(hf)}/hole"]), (tag, hf)
        # numpy in -> numpy out, the way scripts/face_swap.py:253 calls it
        res_np, hole_np = swap_head_mask_revisit_considerGlass(s, t)
        assert isinstance(res_np, np.ndarray) and res_np.dtype == t.dtype
        assert np.array_equal(res_np, mgold[f"swap/{tag}/hair1/res"]) and np.array_equal(hole_np, mgold[f"swap/{tag}/hair1/hole"])
        _, _, fg = swap_head_mask_with_foreground(cu(s), cu(t))
        assert np.array_equal(fg.cpu().numpy(), mgold[f"fg/{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1,), (15,), (16,), (17,), (3, 5, 7), (2, 1, 1024, 1024), (1, 513, 257)])
def test_swap_head_mask_kernel_vs_oracle(shape):
    """Every (source, target) label pair occurs; ragged sizes exercise the 128-bit path, its scalar tail and
    unaligned views; labels above 11 (never produced by the parser) behave like the reference's comparisons."""
    from e4s_b200 import kernels as K
    g = np.random.default_rng(sum(shape))
    s = g.integers(0, 14, size=shape).astype(np.uint8)
    t = g.integers(0, 14, size=shape).astype(np.uint8)
    for hf in (True, False):
        ref_res, ref_hole = MO.swap_head_mask(s, t, hf)
        res, hole, fg = K.swap_head_mask(cu(s), cu(t), hf)
        assert np.array_equal(res.cpu().numpy(), ref_res) and np.array_equal(hole.cpu().numpy(), ref_hole)
        assert np.array_equal(fg.cpu().numpy(), MO.foreground_mask(ref_res, ref_hole))
    if s.size > 40:                                   # unaligned views: offset by 3 bytes -> scalar path
        flat_s, flat_t = cu(s).flatten()[3:], cu(t).flatten()[3:]
        res, hole, _ = K.swap_head_mask(flat_s.clone(), flat_t.clone())
        r2, h2 = MO.swap_head_mask(s.ravel()[3:], t.ravel()[3:])
        assert np.array_equal(res.cpu().numpy(), r2) and np.array_equal(hole.cpu().numpy(), h2)


@pytest.mark.gpu
def test_swap_head_mask_empty_and_errors():
    from e4s_b200 import kernels as K
    e = torch.empty(0, dtype=torch.uint8, device="cuda")
    res, hole, fg = K.swap_head_mask(e, e)
    assert res.numel() == 0 and hole.numel() == 0 and fg.numel() == 0
    with pytest.raises(RuntimeError, match="differ in shape"):
        K.swap_head_mask(torch.zeros(4, dtype=torch.uint8, device="cuda"), torch.zeros(5, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        K.swap_head_mask(torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8))


@pytest.mark.gpu
def test_create_masks_golden(golden, mgold):
    from e4s_b200.masks import create_masks
    for tag in _cases(golden):
        fg = cu(mgold[f"fg/{tag}"]).float()[None, None]
        for r in (0, 1, 5):
            for op in ("dilation", "erosion", "expansion"):
                content, border, full = create_masks(fg, outer_dilation=r, operation=op)
                assert content is fg and border.dtype == torch.float32 and border.shape == fg.shape
                assert np.array_equal(border[0, 0].cpu().numpy(), mgold[f"masks/{tag}/r{r}/{op}/border"].astype(np.float32)), (tag, r, op)
                assert np.array_equal(full[0, 0].cpu().numpy(), mgold[f"masks/{tag}/r{r}/{op}/full"].astype(np.float32)), (tag, r, op)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,radius", [((1, 1, 1, 1), 3), ((2, 3, 31, 65), 0), ((2, 3, 31, 65), 2), ((1, 1, 33, 64), 16),
                                          ((1, 2, 100, 7), 5), ((1, 1, 1024, 1024), 5), ((4, 1, 64, 129), 8)])
def test_box_morph_vs_oracle(shape, radius):
    """uint8 and fp32 images (arbitrary values, not only 0/1), tiles with ragged edges, radius up to the kernel's limit,
    images smaller than the window."""
    from e4s_b200 import kernels as K
    g = np.random.default_rng(radius + shape[-1])
    u8 = (g.random(size=shape) < 0.3).astype(np.uint8) * g.integers(1, 255, size=shape).astype(np.uint8)
    f32 = g.standard_normal(size=shape).astype(np.float32)
    for erode in (False, True):
        ref = MO.box_erosion if erode else MO.box_dilation
        assert np.array_equal(K.mask_box_morph(cu(u8), radius, erode).cpu().numpy(), ref(u8, radius))
        assert np.array_equal(K.mask_box_morph(cu(f32), radius, erode).cpu().numpy(), ref(f32, radius))
    # dilation and erosion are dual under complement; opening is anti-extensive, closing extensive
    b = cu((u8 > 0).astype(np.uint8))
    d, e = K.mask_box_morph(b, radius, False), K.mask_box_morph(b, radius, True)
    assert torch.equal(1 - K.mask_box_morph(1 - b, radius, True), d)
    assert bool((e <= b).all()) and bool((b <= d).all())


@pytest.mark.gpu
def test_morphology_api_errors_and_edge_cases():
    from e4s_b200.masks import dilation, erosion
    x = torch.zeros(1, 1, 8, 8, device="cuda")
    assert torch.equal(dilation(x, torch.ones(3, 3, device="cuda")), x)
    ones = torch.ones(1, 1, 8, 8, device="cuda")
    assert torch.equal(erosion(ones, torch.ones(5, 5, device="cuda")), ones)          # geodesic border: the frame does not erode
    with pytest.raises(ValueError, match="4 dimensions"):
        dilation(torch.zeros(8, 8, device="cuda"), torch.ones(3, 3))
    with pytest.raises(TypeError):
        dilation(np.zeros((1, 1, 8, 8)), torch.ones(3, 3))
    with pytest.raises(NotImplementedError):
        dilation(x, torch.ones(3, 5))
    with pytest.raises(NotImplementedError):
        erosion(x, torch.tensor([[0., 1, 0], [1, 1, 1], [0, 1, 0]]))
    with pytest.raises(RuntimeError, match="E4S_ERR_SHAPE"):
        dilation(x, torch.ones(35, 35))
    assert dilation(torch.zeros(0, 1, 8, 8, device="cuda"), torch.ones(3, 3)).shape == (0, 1, 8, 8)


@pytest.mark.gpu
def test_swap_comp_style_vector_golden(mgold):
    from e4s_b200.masks import swap_comp_style_vector
    comp = sorted(set(range(12)) - {0, 4, 11, 10})
    for tag in ("plain", "no_ear", "no_teeth", "neither"):
        for interp in (0, 1):
            k = f"stylevec/{tag}/interp{interp}"
            out = swap_comp_style_vector(cu(mgold[k + "/sv1"]), cu(mgold[k + "/sv2"]), comp, belowFace_interpolation=bool(interp))
            assert np.array_equal(out.cpu().numpy(), mgold[k + "/out"]), k


@pytest.mark.gpu
def test_swapped_mask_drives_the_generator(golden):
    """Step 4 -> step 5 of scripts/face_swap.py on the device: the swapped label map of the two example faces, one-hot
    encoded, is a valid region mask for the generator (no host round trip in between)."""
    from e4s_b200.masks import labelMap2OneHot, swap_head_mask_revisit_considerGlass
    from e4s_b200.stylegan2.modconv import LabelPyramid
    s, t = _cases(golden)["example"]
    res, _ = swap_head_mask_revisit_considerGlass(cu(s), cu(t))
    onehot = labelMap2OneHot(res[None, None], 12)
    pyr = LabelPyramid.from_mask(onehot)
    assert torch.equal(pyr.at(*res.shape), res[None])


@pytest.mark.gpu
def test_swap_faces_pipeline_matches_oracle_composition(golden):
    """e4s_b200.face_swap.swap_faces (steps 3-5 + blending masks of scripts/face_swap.py, batched on the device) against
    the same steps composed on the CPU from the oracle's restatements: encoder -> shape swap -> texture swap -> MLPs ->
    generator.  Labels / masks bit-exact, images within the fp32 bar."""
    import types
    from conftest import REL_TOL, assert_close
    from oracle import e4s_oracle as O
    from e4s_b200.face_swap import swap_faces
    from e4s_b200.networks import Net3
    opts = types.SimpleNamespace(fsencoder_type="psp", remaining_layer_idx=13, num_seg_cls=12, out_size=64,
                                 train_G=False, start_from_latent_avg=True, learn_in_w=False)
    net = Net3(opts).eval()
    st = O.synthetic_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, salt=5)
    net.load_state_dict(st)
    net = net.cuda()
    g = torch.Generator().manual_seed(21)
    lat = 0.1 * torch.randn(18, 512, generator=g)
    net.latent_avg = lat.cuda()
    rnd = synthetic_label_maps(31, 2, 128, 128)
    d_lab, t_lab = torch.from_numpy(rnd[:1].copy()), torch.from_numpy(rnd[1:].copy())
    d_lab[(d_lab == 7) | (d_lab == 9)] = 6                  # a source without ears and teeth: both special cases of the texture swap
    driven, target = torch.randn(1, 3, 256, 256, generator=g), torch.randn(1, 3, 256, 256, generator=g)
    _, _, _, noise = O.synthetic_inputs(1, 12, 64, 128, seed=9)
    out = swap_faces(net, driven.cuda(), target.cuda(), d_lab.cuda(), t_lab.cuda(), noise=[n.cuda() for n in noise])
    # ---- the same on the CPU
    d_vec, _ = O.get_style_vectors(st, driven, O.label_to_onehot(d_lab[:, None].long(), 12))
    t_vec, _ = O.get_style_vectors(st, target, O.label_to_onehot(t_lab[:, None].long(), 12))
    res, hole = MO.swap_head_mask(d_lab[0].numpy(), t_lab[0].numpy())
    comp = sorted(set(range(12)) - {0, 4, 11, 10})
    vec = torch.from_numpy(MO.swap_comp_style_vector(t_vec.numpy(), d_vec.numpy(), comp))
    assert float(d_vec[:, 9].abs().sum()) == 0.0 and float(d_vec[:, 7].abs().sum()) == 0.0     # both branches are taken
    codes = O.cal_style_codes(st, vec, lat, 13)
    gst = {k[2:]: v for k, v in st.items() if k.startswith("G.")}
    ref_img, _ = O.generator_forward(gst, codes, O.label_to_onehot(torch.from_numpy(res)[None, None].long(), 12), noise, 64, 13)
    fg = MO.foreground_mask(res, hole)
    _, border, full = MO.create_masks(fg, 5, "dilation")
    assert np.array_equal(out.swapped_label[0].cpu().numpy(), res) and np.array_equal(out.hole_map[0].cpu().numpy(), hole)
    assert np.array_equal(out.content_mask[0, 0].cpu().numpy(), fg.astype(np.float32))
    assert np.array_equal(out.border_mask[0, 0].cpu().numpy(), border.astype(np.float32))
    assert np.array_equal(out.full_mask[0, 0].cpu().numpy(), full.astype(np.float32))
    assert_close(out.style_vectors, vec, REL_TOL, "swapped texture vectors")
    assert_close(out.image, ref_img, REL_TOL, "swapped face")
