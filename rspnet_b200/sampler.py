"""On-GPU clip sampling for the pretraining loader (subsystem 4 of the hot path).

Host side: the reference's random decisions, drawn from python's ``random`` in the reference's order
(SURVEY.md appendix B): per video ``RandomStrideCrop`` twice (two clips; datasets/transforms_video/
transforms_temporal.py:25-50, fallback :16-22), then per clip ``RawVideoRandomCrop.get_params``
(transforms_spatial.py:50-83), then per clip on the main process ``RandomGrayScale`` and
``ColorJitter`` (four ``random.uniform`` factors then ``random.shuffle`` of the op list, transforms_tensor.py:97-125) and
``RandomHorizontalFlipVideo`` (transforms_tensor.py:13-31; torchvision) — the order of the non-``aug_plus`` chain
of datasets/classification/__init__.py:188-202.  Device side: ``rsp_clip_sample`` / ``rsp_clip_sample_jitter``
replace the per-clip loop of ``SequentialGPUCollateFn`` (transforms_tensor.py:214-233): uint8 frame gather + crop +
bilinear resize + gray + colour jitter + flip + normalise for the whole batch in one kernel (plus one reduction
pass for the clip-wide gray mean the contrast op blends with).

The ``aug_plus`` chain (RandomApply(ColorJitter) / GaussianBlur) is not implemented.
"""
import ctypes as C
import math
import random
from bisect import bisect_left
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def calc_needed_frames(size: int, stride: int) -> int:
    return (size - 1) * stride + 1


def fallback_select(size: int, stride: int, num_frames: int):
    """Short videos: wrap around, or spread evenly (transforms_temporal.py:16-22)."""
    assert num_frames > 0, 'No frames in video'
    if num_frames <= size:
        return np.arange(size) % num_frames
    if num_frames < calc_needed_frames(size, stride):
        return np.linspace(0, num_frames - 1, num=size).round().astype(int)
    return None


class RandomStrideCrop:
    """``size`` frame indices at a stride drawn from a weighted list; draws: random.random() then random.randint()."""

    def __init__(self, size: int, strides=({'stride': 1, 'weight': 1},)):
        self.size = size
        self.set_strides(strides)

    def set_strides(self, strides=({'stride': 1, 'weight': 1},)):
        self.strides = [dict(s) for s in strides]
        total = sum(s['weight'] for s in self.strides)
        acc, self.prefix_weight_sum = 0, []
        for s in self.strides:
            s['weight'] /= total
            acc += s['weight']
            self.prefix_weight_sum.append(acc)

    def set_size(self, size: int):
        self.size = size

    def __call__(self, frame_indices: np.ndarray) -> np.ndarray:
        num_frames = len(frame_indices)
        stride = self.strides[bisect_left(self.prefix_weight_sum, random.random())]['stride']
        selected = fallback_select(self.size, stride, num_frames)
        if selected is None:
            needed = calc_needed_frames(self.size, stride)
            start = random.randint(0, num_frames - needed)
            selected = np.arange(start, start + needed, stride)
        return frame_indices[selected]


class RawVideoRandomCrop:
    """Random-resized-crop box on the raw frame (scale / log-uniform ratio, 10 tries, centre fallback)."""

    def __init__(self, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.)):
        self.scale = scale
        self.ratio = ratio

    def get_params(self, height: int, width: int) -> Tuple[int, int, int, int]:
        area = height * width
        for _ in range(10):
            target_area = random.uniform(*self.scale) * area
            log_ratio = (math.log(self.ratio[0]), math.log(self.ratio[1]))
            aspect_ratio = math.exp(random.uniform(*log_ratio))
            w = int(round(math.sqrt(target_area * aspect_ratio)))
            h = int(round(math.sqrt(target_area / aspect_ratio)))
            if 0 < w <= width and 0 < h <= height:
                i = random.randint(0, height - h)
                j = random.randint(0, width - w)
                return i, j, h, w
        in_ratio = float(width) / float(height)
        if in_ratio < min(self.ratio):
            w = width
            h = int(round(w / min(self.ratio)))
        elif in_ratio > max(self.ratio):
            h = height
            w = int(round(h * max(self.ratio)))
        else:
            w, h = width, height
        return (height - h) // 2, (width - w) // 2, h, w


JITTER_DTYPE = np.dtype([("factor", np.float32, (4,)), ("order", np.uint8, (4,))])   # struct ClipJitter of the C ABI
JITTER_OPS = ("brightness", "contrast", "saturation", "hue")


class ColorJitter:
    """Random decisions of the reference's ColorJitter (transforms_tensor.py:54-145): ranges as ``_check_input`` builds
    them, one ``random.uniform`` per enabled op in the fixed order brightness, contrast, saturation, hue, then
    ``random.shuffle`` of the enabled ops."""

    def __init__(self, brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0):
        self.ranges = [self._range(brightness, "brightness"), self._range(contrast, "contrast"),
                       self._range(saturation, "saturation"),
                       self._range(hue, "hue", center=0, bound=(-0.5, 0.5), clip_first_on_zero=False)]

    @staticmethod
    def _range(value, name, center=1, bound=(0, float("inf")), clip_first_on_zero=True):
        if isinstance(value, (int, float)):
            if value < 0:
                raise ValueError("If {} is a single number, it must be non negative.".format(name))
            value = [center - value, center + value]
            if clip_first_on_zero:
                value[0] = max(value[0], 0)
        elif isinstance(value, (tuple, list)) and len(value) == 2:
            if not bound[0] <= value[0] <= value[1] <= bound[1]:
                raise ValueError("{} values should be between {}".format(name, bound))
        else:
            raise TypeError("{} should be a single number or a list/tuple with lenght 2.".format(name))
        return None if value[0] == value[1] == center else value

    def get_params(self):
        """Returns (factor[4], order[4]) — op ids in application order, 255-padded."""
        factor, ops = [0.0] * 4, []
        for op, rng in enumerate(self.ranges):
            if rng is not None:
                factor[op] = random.uniform(rng[0], rng[1])
                ops.append(op)
        random.shuffle(ops)
        if not -0.5 <= factor[3] <= 0.5:
            raise ValueError("hue_factor is not in [-0.5, 0.5].")
        return factor, ops + [255] * (4 - len(ops))


def clip_sample(frames: torch.Tensor, frame_idx: torch.Tensor, boxes: torch.Tensor, flags: torch.Tensor,
                mean: Sequence[float], std: Sequence[float], size: int, layout: int = 0,
                jitter: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frames uint8 [F,H,W,3] (device); frame_idx int32 [n,T]; boxes int32 [n,4]; flags uint8 [n]; jitter: uint8
    [n, 20] device view of a ``JITTER_DTYPE`` table, or None.  ``out``: an existing contiguous clip tensor of the
    layout's shape and dtype to fill in place (e.g. the input buffers of a captured training step)."""
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.shape[-1] == 3
    n, t = frame_idx.shape
    _, hs, ws, _ = frames.shape
    dev = frames.device
    shape, dtype = ((n, 3, t, size, size), torch.float32) if layout == 0 else ((n, t, size, size, 4), torch.bfloat16)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=dev)
    elif tuple(out.shape) != shape or out.dtype != dtype or out.device != dev or not out.is_contiguous():
        raise ValueError(f"clip_sample: out must be a contiguous {dtype} tensor of shape {shape} on {dev}")
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    if jitter is None:
        call("rsp_clip_sample", ptr(frames), ptr(frame_idx), ptr(boxes), ptr(flags), m3, s3, n, t, hs, ws, size,
             layout, ptr(out), stream_ptr())
    else:
        assert jitter.dtype == torch.uint8 and jitter.shape == (n, JITTER_DTYPE.itemsize) and jitter.is_contiguous()
        sums = torch.empty(n, dtype=torch.float32, device=dev)
        call("rsp_clip_sample_jitter", ptr(frames), ptr(frame_idx), ptr(boxes), ptr(flags), ptr(jitter), ptr(sums),
             m3, s3, n, t, hs, ws, size, layout, ptr(out), stream_ptr())
    return out


def jitter_table(records) -> torch.Tensor:
    """[(factor[4], order[4]), ...] -> host uint8 [n, 20] tensor laid out as ``struct ClipJitter``."""
    tab = np.zeros(len(records), dtype=JITTER_DTYPE)
    if records:
        tab["factor"] = np.asarray([r[0] for r in records], dtype=np.float32)
        tab["order"] = np.asarray([r[1] for r in records], dtype=np.uint8)
    return torch.from_numpy(tab.view(np.uint8).reshape(len(records), JITTER_DTYPE.itemsize))


class GPUClipSampler:
    """Produces the loader contract ``((clip_q, clip_k), None)`` (datasets/classification/__init__.py:22-50) from a
    device-resident pool of decoded uint8 frames: ``frames`` [F,H,W,3] with ``video_offsets``/``video_lengths``
    describing where each video's frames live."""

    def __init__(self, size: int = 112, temporal_size: int = 32, strides=({'stride': 1, 'weight': 1},),
                 crop_scale=(0.4, 1.0), gray_p: float = 0.2, flip_p: float = 0.5,
                 mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), color_jitter=None):
        """``color_jitter``: None (the ``no_color_jitter`` configs) or a dict / 4-tuple of ColorJitter arguments —
        the reference's pretraining chain uses ``dict(brightness=.4, contrast=.4, saturation=.4, hue=.4)``."""
        if isinstance(color_jitter, dict):
            color_jitter = ColorJitter(**color_jitter)
        elif isinstance(color_jitter, (tuple, list)):
            color_jitter = ColorJitter(*color_jitter)
        self.jitter = color_jitter
        self.size = size
        self.temporal = RandomStrideCrop(temporal_size, strides)
        self.crop = RawVideoRandomCrop(scale=crop_scale)
        self.gray_p, self.flip_p = gray_p, flip_p
        self.mean, self.std = list(mean), list(std)

    def draw(self, video_lengths: Sequence[int], height: int, width: int):
        """All random decisions for one batch, in the reference's order. Returns numpy arrays in clip-major layout
        [2][B]: frame indices (video-relative), boxes, flags, and the colour-jitter table (None without jitter)."""
        b = len(video_lengths)
        t = self.temporal.size
        idx = np.zeros((2, b, t), dtype=np.int32)
        box = np.zeros((2, b, 4), dtype=np.int32)
        flags = np.zeros((2, b), dtype=np.uint8)
        bases = {}
        for v, n_frames in enumerate(video_lengths):          # worker side: per video
            base = bases.get(n_frames)
            if base is None:
                base = bases[n_frames] = np.arange(n_frames)
            for c in range(2):
                idx[c, v] = self.temporal(base)
            for c in range(2):
                box[c, v] = self.crop.get_params(height, width)
        jit = [[None] * b, [None] * b] if self.jitter is not None else None
        for v in range(b):                                     # main process: per video, per clip GPU transforms
            for c in range(2):
                gray = random.random() < self.gray_p
                if jit is not None:
                    jit[c][v] = self.jitter.get_params()
                flip = random.random() < self.flip_p
                flags[c, v] = (1 if flip else 0) | (2 if gray else 0)
        return idx, box, flags, (jitter_table(jit[0] + jit[1]) if jit is not None else None)

    def __call__(self, frames: torch.Tensor, video_offsets: Sequence[int], video_lengths: Sequence[int],
                 layout: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """``out``: optional ``(clip_q, clip_k)`` buffers to fill in place — ``PretrainEngine.next_input_pair()`` hands
        out the pair the next captured step reads, which saves the copy of both clips into the graph's inputs."""
        _, h, w, _ = frames.shape
        idx, box, flags, jit = self.draw(video_lengths, h, w)
        b = len(video_lengths)
        idx = idx + np.asarray(video_offsets, dtype=np.int32)[None, :, None]
        dev = frames.device
        # all per-clip tables travel in ONE pinned staging buffer and one async copy: a pageable source would make the
        # host wait for the stream to reach the copy, i.e. for the previous training step, on every batch
        n, t = 2 * b, idx.shape[-1]
        o_box, o_jit = n * t * 4, n * t * 4 + n * 16
        o_flags = o_jit + n * JITTER_DTYPE.itemsize
        stage = torch.empty(o_flags + n, dtype=torch.uint8, pin_memory=dev.type == "cuda")
        hv = stage.numpy()
        hv[:o_box].view(np.int32)[:] = idx.reshape(-1)
        hv[o_box:o_jit].view(np.int32)[:] = box.reshape(-1)
        if jit is not None:
            hv[o_jit:o_flags] = jit.numpy().reshape(-1)
        hv[o_flags:] = flags.reshape(-1)
        dbuf = stage.to(dev, non_blocking=True)
        t_idx = dbuf[:o_box].view(torch.int32).view(n, t)
        t_box = dbuf[o_box:o_jit].view(torch.int32).view(n, 4)
        t_jit = dbuf[o_jit:o_flags].view(n, JITTER_DTYPE.itemsize) if jit is not None else None
        t_flags = dbuf[o_flags:]
        if out is None:
            out = clip_sample(frames, t_idx, t_box, t_flags, self.mean, self.std, self.size, layout, jitter=t_jit)
            return (out[:b], out[b:]), None
        for c, dst in enumerate(out):   # caller-owned buffers need not be adjacent: one launch per clip set
            rows = slice(c * b, (c + 1) * b)
            clip_sample(frames, t_idx[rows], t_box[rows], t_flags[rows], self.mean, self.std, self.size, layout,
                        jitter=None if t_jit is None else t_jit[rows], out=dst)
        return (out[0], out[1]), None
