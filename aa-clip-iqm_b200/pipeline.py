"""Batch-level public API of the hot path: images in, anomaly maps + image scores out.

Mirrors one iteration of the reference's evaluation loop, test_last.get_predictions (test_last.py:70-151, no-IQM
branch): ``image.to(device)`` -> ``model(image)`` -> ``det @ text`` -> 4 x ``calculate_similarity_map(test=True)`` ->
level sum -> ``.cpu().numpy()``. Here the device part is two C-ABI calls (aaclip_forward, aaclip_score_maps) and the
host<->device copies go through pinned staging buffers on the current stream.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import engine as _engine

DOMAIN_BLUR = {"Industrial": (7, 1.0), "Medical": (9, 1.5)}  # forward_utils.py:205-206


class AnomalyPredictor:
    """Adapted ViT-L-14-336 tower + patch-text scoring on one GPU."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], text: torch.Tensor, img_size: int = 518,
                 device: str | torch.device = "cuda", domain: str = "Industrial", max_batch: int = 32,
                 relu: bool = False, **engine_cfg):
        if domain not in DOMAIN_BLUR:
            raise ValueError(f"domain must be one of {sorted(DOMAIN_BLUR)}")
        self.device = torch.device(device)
        self.img_size = img_size
        self.domain = domain
        self.max_batch = max_batch
        self.engine = _engine.AaclipEngine(self.device, img_size=img_size, max_batch=max_batch,
                                           proj_relu=int(relu), **engine_cfg)
        self.engine.load_state_dict(state_dict)
        if self.engine.missing:
            raise RuntimeError(f"{self.engine.missing} weight tensors of the visual path are missing from the state dict")
        # the patch features go straight into the scoring call: skip the stand-alone F.normalize passes over them
        # (model/adapter.py:182) and let the similarity kernel normalise in the pass it makes anyway
        self.engine.seg_normalize(False)
        self.text = text.to(self.device, torch.float32).contiguous()
        self._stage_in: Optional[torch.Tensor] = None
        self._stage_guard = _engine.StreamGuard()       # _stage_in is shared by predict_u8 / predict_host calls
        self._pre = None

    # -- device-resident entry point (what bench.py's `value` times)
    def predict_device(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """images [B,3,S,S] fp32 already on the GPU -> (maps [B,S,S], scores [B]) on the GPU."""
        seg, det, _, _ = self.engine.forward(images)
        k, s = DOMAIN_BLUR[self.domain]
        return _engine.score_maps(seg, self.text, self.img_size, test=True, ksize=k, sigma=s, det=det, raw=True)

    # -- the same step as a CUDA graph (one launch per batch instead of 157)
    def capture(self, batch: int) -> "CapturedStep":
        """Captures predict_device for `batch` images into a CUDA graph with static input / output buffers. The C ABI
        allocates nothing and never synchronises after aaclip_finalize_weights, so the whole step -- tensor-map
        encoding happens on the host at capture time -- replays from one cudaGraphLaunch. Worth it when the HOST is the
        bottleneck (small batches, small images); at B = 32 @518 the GPU is never waiting for launches (bench.py --graph)."""
        return CapturedStep(self, batch)

    # -- decoded images in: the dataset transform runs on the device too
    def predict_u8(self, images_u8) -> Tuple[torch.Tensor, torch.Tensor]:
        """images [B,H,W,3] uint8 (decoded RGB, one size) -> (maps [B,S,S], scores [B]) on the GPU. The reference's
        per-image CPU transform (dataset/__init__.py:150-159: bicubic resize, ToTensor, Normalize) is executed by
        aaclip_preprocess_u8, bit-exact with PIL + torchvision, straight into the tower's input layout."""
        if self._pre is None:
            from . import preprocess
            self._pre = preprocess.ImagePreprocessor(self.img_size, device=self.device)
        x = torch.as_tensor(images_u8)
        B = x.shape[0]
        if self._stage_in is None or self._stage_in.shape[0] < B:
            self._stage_in = torch.empty((max(B, self.max_batch), 3, self.img_size, self.img_size), dtype=torch.float32,
                                         device=self.device)
        self._stage_guard.enter()
        out = self.predict_device(self._pre(x, out=self._stage_in[:B]))
        self._stage_guard.leave()
        return out

    @property
    def launches_per_batch(self) -> int:
        """Kernel launches of one predict_device call (tower + similarity + map + image-score kernels)."""
        return self.engine.last_launch_count + 3

    # -- host entry point (what bench.py's `e2e` times)
    def predict_host(self, images_pinned: torch.Tensor, maps_out_pinned: torch.Tensor,
                     scores_out_pinned: torch.Tensor) -> None:
        """Pinned host images -> pinned host maps/scores; H2D, compute and D2H are enqueued on the current stream.
        The caller synchronises (or records an event) before reading the outputs."""
        B = images_pinned.shape[0]
        if self._stage_in is None or self._stage_in.shape[0] < B:
            self._stage_in = torch.empty((max(B, self.max_batch), 3, self.img_size, self.img_size), dtype=torch.float32,
                                         device=self.device)
        dev_in = self._stage_in[:B]
        self._stage_guard.enter()
        dev_in.copy_(images_pinned, non_blocking=True)
        maps, scores = self.predict_device(dev_in)
        self._stage_guard.leave()
        maps_out_pinned.copy_(maps, non_blocking=True)
        scores_out_pinned.copy_(scores, non_blocking=True)


class CapturedStep:
    """predict_device as a replayable CUDA graph: ``step = predictor.capture(B); maps, scores = step(images)``.
    `images` are copied into the graph's static input; the returned tensors are the graph's static outputs (valid until
    the next call)."""

    def __init__(self, predictor: AnomalyPredictor, batch: int):
        p = predictor
        self.static_in = torch.zeros((batch, 3, p.img_size, p.img_size), dtype=torch.float32, device=p.device)
        side = torch.cuda.Stream(device=p.device)
        side.wait_stream(torch.cuda.current_stream(p.device))
        with torch.cuda.stream(side):             # warm-up outside the capture: function attributes, workspaces
            for _ in range(2):
                p.predict_device(self.static_in)
        torch.cuda.current_stream(p.device).wait_stream(side)
        torch.cuda.synchronize(p.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.maps, self.scores = p.predict_device(self.static_in)

    def __call__(self, images: torch.Tensor):
        if images.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(images, non_blocking=True)
        self.graph.replay()
        return self.maps, self.scores


class HostStream:
    """Pipelined host-to-host evaluation: pinned host images in, pinned host maps / scores out, with the PCIe copies of
    batch i+1 and batch i-1 overlapped with the kernels of batch i (separate copy streams, two device slots).

    This is the steady state of the reference's evaluation loop (test_last.py:70-151: ``image.to(device)`` ...
    ``final_map.cpu().numpy()`` per batch) without its per-batch synchronisation. ``submit`` returns a CUDA event;
    the outputs of that batch are valid once the event has completed (``event.synchronize()``)."""

    def __init__(self, predictor: AnomalyPredictor, slots: int = 2):
        self.p = predictor
        dev, B, S = predictor.device, predictor.max_batch, predictor.img_size
        self.slots = slots
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.compute = torch.cuda.Stream(device=dev)
        self.dev_in = [torch.empty((B, 3, S, S), dtype=torch.float32, device=dev) for _ in range(slots)]
        self.dev_maps = [None] * slots          # keeps the device outputs of a slot alive until its D2H has run
        self.in_free = [torch.cuda.Event() for _ in range(slots)]     # compute has consumed dev_in[slot]
        self.out_done = [torch.cuda.Event() for _ in range(slots)]    # D2H of the slot's outputs has completed
        self.i = 0

    def submit(self, images_pinned: torch.Tensor, maps_out_pinned: torch.Tensor,
               scores_out_pinned: torch.Tensor) -> torch.cuda.Event:
        slot = self.i % self.slots
        first_use = self.i < self.slots
        self.i += 1
        B = images_pinned.shape[0]
        with torch.cuda.stream(self.h2d):
            if not first_use:
                self.h2d.wait_event(self.in_free[slot])            # the previous occupant has been read by the kernels
            self.dev_in[slot][:B].copy_(images_pinned, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.h2d)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(copied)
            if not first_use:
                self.compute.wait_event(self.out_done[slot])       # the slot's previous outputs have left the device
            maps, scores = self.p.predict_device(self.dev_in[slot][:B])
            self.in_free[slot].record(self.compute)
            computed = torch.cuda.Event()
            computed.record(self.compute)
        self.dev_maps[slot] = (maps, scores)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(computed)
            maps_out_pinned.copy_(maps, non_blocking=True)
            scores_out_pinned.copy_(scores, non_blocking=True)
            maps.record_stream(self.d2h)         # allocated on the compute stream, read here
            scores.record_stream(self.d2h)
            self.out_done[slot].record(self.d2h)
        return self.out_done[slot]

    def drain(self):
        self.d2h.synchronize()
        self.compute.synchronize()
