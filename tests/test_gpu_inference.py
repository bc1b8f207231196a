"""GPU parity of the inference wrappers against the UNMODIFIED reference functions
(seggpt_engine.run_one_image / inference_image / inference_video, painter_inference_segm.run_one_image) on the
reference's own example images and videos, with seeded random weights.

Protocol: both sides are driven with the SAME painter_b200 module (it is a drop-in for the reference functions), so
the forward numerics are identical and any difference is in the wrapper itself: the device kernels for stitch /
normalise / layout / de-normalise / resize / blend must reproduce the reference's numpy / torch-CPU arithmetic bit for
bit (uint8 files identical, float64 tensors identical)."""
import os

import numpy as np
import pytest
import torch

from oracle import painter_oracle as po
from oracle import ref_loader

from _common import build_model
from _refmods import have_reference

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_reference(), reason="reference tree not staged")]


def _ex(name):
    return os.path.join(ref_loader.examples_dir(), name)


@pytest.fixture(scope="module")
def seggpt():
    cfg = po.PainterConfig(seggpt=True)
    model, _ = build_model(cfg, 3)
    model.eval()
    model.seg_type = "instance"
    return model


def test_run_one_image_matches_reference_function(seggpt):
    from painter_b200 import seggpt_engine as ours
    ref = ref_loader.seggpt_engine()
    dev = torch.device("cuda")
    rng = np.random.RandomState(0)
    for P in (1, 2):
        img = rng.randn(P, 896, 448, 3)
        tgt = rng.randn(P, 896, 448, 3)
        a = ref.run_one_image(img, tgt, seggpt, dev)
        b = ours.run_one_image(img, tgt, seggpt, dev)
        assert a.dtype == b.dtype == torch.float64 and a.shape == b.shape == (448, 448, 3)
        assert torch.equal(a, b), (a - b).abs().max()


def test_inference_image_writes_the_same_png(seggpt, tmp_path):
    from PIL import Image
    from painter_b200 import seggpt_engine as ours
    ref = ref_loader.seggpt_engine()
    dev = torch.device("cuda")
    cases = [("hmbb_3.jpg", ["hmbb_1.jpg"], ["hmbb_1_target.png"]),
             ("hmbb_3.jpg", ["hmbb_1.jpg", "hmbb_2.jpg"], ["hmbb_1_target.png", "hmbb_2_target.png"])]
    for k, (q, ps, ts) in enumerate(cases):
        o_ref, o_our = str(tmp_path / f"ref{k}.png"), str(tmp_path / f"our{k}.png")
        ref.inference_image(seggpt, dev, _ex(q), [_ex(p) for p in ps], [_ex(t) for t in ts], o_ref)
        ours.inference_image(seggpt, dev, _ex(q), [_ex(p) for p in ps], [_ex(t) for t in ts], o_our)
        a, b = np.array(Image.open(o_ref)), np.array(Image.open(o_our))
        assert a.shape == b.shape and np.array_equal(a, b), np.abs(a.astype(int) - b.astype(int)).max()


def test_inference_video_rolling_prompt_cache_same_frames(seggpt, tmp_path):
    import cv2
    from painter_b200 import seggpt_engine as ours
    ref = ref_loader.seggpt_engine()
    dev = torch.device("cuda")
    for num_frames in (0, 2):
        o_ref, o_our = str(tmp_path / f"ref{num_frames}.mp4"), str(tmp_path / f"our{num_frames}.mp4")
        ref.inference_video(seggpt, dev, _ex("video_1.mp4"), num_frames, [_ex("video_1.jpg")],
                            [_ex("video_1_target.png")], o_ref)
        n = ours.inference_video(seggpt, dev, _ex("video_1.mp4"), num_frames, [_ex("video_1.jpg")],
                                 [_ex("video_1_target.png")], o_our)
        ca, cb = cv2.VideoCapture(o_ref), cv2.VideoCapture(o_our)
        frames = 0
        while True:
            ra, fa = ca.read()
            rb, fb = cb.read()
            assert ra == rb
            if not ra:
                break
            assert np.array_equal(fa, fb), (num_frames, frames, np.abs(fa.astype(int) - fb.astype(int)).max())
            frames += 1
        assert frames == n == 8


def test_painter_run_one_image_writes_the_same_png(tmp_path):
    from PIL import Image
    from painter_b200 import painter_inference as ours
    ref = ref_loader.painter_inference_segm()
    cfg = po.PainterConfig()
    model, _ = build_model(cfg, 1)
    model.eval()

    class Wrapped(torch.nn.Module):      # the script calls model(...) and model.module.* (a DDP wrapper)
        def __init__(self, m):
            super().__init__()
            self.module = m

        def forward(self, *a, **k):
            return self.module(*a, **k)

    w = Wrapped(model)
    dev = torch.device("cuda")
    rng = np.random.RandomState(1)
    img, tgt = rng.randn(896, 448, 3), rng.randn(896, 448, 3)
    for size in ((700, 400), (448, 448), (301, 523)):
        o_ref, o_our = str(tmp_path / "ref.png"), str(tmp_path / "our.png")
        with torch.no_grad():
            ref.run_one_image(img, tgt, size, w, o_ref, dev)
        ours.run_one_image(img, tgt, size, w, o_our, dev)
        a, b = np.array(Image.open(o_ref)), np.array(Image.open(o_our))
        assert a.shape == b.shape == (size[1], size[0], 3)
        assert np.array_equal(a, b), (size, np.abs(a.astype(int) - b.astype(int)).max())
