"""Clip pipeline (subsystem 4): host-side random decisions bit-exact against fixtures recorded from the reference
transforms; the fused GPU kernel against the reference's per-clip transform chain."""
import json
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN


def test_random_decisions_match_reference_fixture():
    from rspnet_b200.sampler import RandomStrideCrop, RawVideoRandomCrop
    rec = json.loads((GOLDEN / "sampler_draws.json").read_text())
    random.seed(1234)
    crops = {}
    crop = RawVideoRandomCrop(scale=(0.4, 1.0))
    last_key, tc = None, None
    for r in rec:
        if r["kind"] == "temporal":
            key = (r["size"], json.dumps(r["strides"]), r["n"])
            if key != last_key:
                tc = RandomStrideCrop(size=r["size"], strides=r["strides"])
                last_key = key
            assert [int(v) for v in tc(np.arange(r["n"]))] == r["out"]
        elif r["kind"] == "crop":
            assert list(crop.get_params(r["h"], r["w"])) == r["out"]
        else:
            assert random.random() == r["value"]     # same number of draws consumed


def test_fallback_select_edges():
    from rspnet_b200.sampler import fallback_select
    assert fallback_select(4, 1, 3).tolist() == [0, 1, 2, 0]
    assert fallback_select(4, 3, 7).tolist() == [0, 2, 4, 6]
    assert fallback_select(4, 1, 10) is None
    with pytest.raises(AssertionError):
        fallback_select(4, 1, 0)


@pytest.mark.gpu
def test_clip_kernel_matches_reference_chain():
    from rspnet_b200 import sampler
    g = torch.load(GOLDEN / "sampler_clip.pt")
    out = sampler.clip_sample(g["frames"].cuda(), g["idx"].cuda(), g["box"].cuda(), g["flags"].cuda(), g["mean"],
                              g["std"], g["size"], layout=0)
    torch.testing.assert_close(out.cpu(), g["out"], rtol=1e-5, atol=2e-5)
    out1 = sampler.clip_sample(g["frames"].cuda(), g["idx"].cuda(), g["box"].cuda(), g["flags"].cuda(), g["mean"],
                               g["std"], g["size"], layout=1)
    exp = g["out"].permute(0, 2, 3, 4, 1).bfloat16()
    assert (out1[..., :3].cpu().float() - exp.float()).abs().max() < 2e-2
    assert torch.count_nonzero(out1[..., 3]) == 0


@pytest.mark.gpu
def test_gpu_clip_sampler_contract():
    from rspnet_b200.sampler import GPUClipSampler
    random.seed(0)
    frames = torch.randint(0, 256, (96 * 3, 64, 80, 3), dtype=torch.uint8, device="cuda")
    s = GPUClipSampler(size=32, temporal_size=8)
    (clip_q, clip_k), label = s(frames, [0, 96, 192], [96, 96, 96])
    assert label is None and clip_q.shape == clip_k.shape == (3, 3, 8, 32, 32) and clip_q.dtype == torch.float32
    assert torch.isfinite(clip_q).all() and torch.isfinite(clip_k).all()
