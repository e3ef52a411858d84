"""On-device evaluation metrics: the AUROC / average-precision part of the reference's
``metrics_eval`` (forward_utils.py:233-308) without copying the anomaly maps to the host and without a CPU sort.

The reference min-max normalises the maps of a whole class, takes the per-image maximum, and hands 268 324 pixels
per image to scikit-learn. Here the maps stay on the GPU (they were produced there), and every statistic that has to
be global -- min, max, two score histograms -- is additive across data-parallel ranks: ``metrics_eval_device`` with a
process group all-reduces a few MB instead of gathering 1.07 MB per image to a sink.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .engine import _need_cuda, _ptr, _stream

N_BINS = 1 << 20   # discretisation error of AUROC / AP <= 2^-21 (pairs sharing a bin)


def _ws(device) -> torch.Tensor:
    return torch.empty(_lib.load().aaclip_metrics_workspace_bytes(), dtype=torch.uint8, device=device)


def minmax(chunks: Sequence[torch.Tensor]) -> torch.Tensor:
    """Device tensor [2] = (min, max) over all fp32 chunks (aaclip_minmax, streaming merge)."""
    lib = _lib.load()
    dev = chunks[0].device
    out = torch.empty(2, dtype=torch.float32, device=dev)
    ws = _ws(dev)
    for i, c in enumerate(chunks):
        _need_cuda(c)
        c = c.contiguous().float()
        with torch.cuda.device(dev):
            _lib.check(lib.aaclip_minmax(_ptr(c), c.numel(), _ptr(ws), _ptr(out), int(i > 0), _stream()), "aaclip_minmax")
    return out


def _multi_rank(group=None) -> bool:
    """True when torch.distributed is initialised with more than one rank in `group` (None = the default group). The
    ONE predicate every reduction / gather below uses, so that pixel-level and image-level statistics are always
    reduced over the same set of ranks."""
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def gather_concat(t: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather 1-D tensors of (possibly) different lengths and concatenate them in rank order on every rank
    (identity in a single process)."""
    if not _multi_rank(group):
        return t
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device=t.device), group=group)
    mx = int(max(s.item() for s in sizes))
    buf = torch.zeros(mx, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    outs = [torch.empty_like(buf) for _ in sizes]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])


def reduce_minmax(mm: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce a (min, max) pair over the data-parallel ranks (identity without a process group)."""
    if _multi_rank(group):
        lo, hi = mm[0:1].clone(), mm[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        return torch.cat([lo, hi])
    return mm


def reduce_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce (sum) the [2, n_bins] count histograms over the ranks. Counts travel as int64."""
    if _multi_rank(group):
        h = hist.to(torch.int64)
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        return h
    return hist


def histogram(chunks: Sequence[torch.Tensor], labels: Sequence[torch.Tensor], mm: Optional[torch.Tensor],
              n_bins: int = N_BINS) -> torch.Tensor:
    """uint32 counts [2, n_bins] (row 0 negatives, row 1 positives) of the (min-max normalised if mm is given) scores."""
    lib = _lib.load()
    dev = chunks[0].device
    hist = torch.zeros((2, n_bins), dtype=torch.int32, device=dev)     # reinterpreted as uint32 by the kernel
    for c, l in zip(chunks, labels):
        _need_cuda(c, l)
        c = c.contiguous().float()
        l = l.contiguous().to(torch.uint8)
        if l.numel() != c.numel():
            raise ValueError(f"labels ({l.numel()}) and scores ({c.numel()}) differ in size")
        with torch.cuda.device(dev):
            _lib.check(lib.aaclip_score_histogram(_ptr(c), _ptr(l), c.numel(), _ptr(mm), n_bins, _ptr(hist), _stream()),
                       "aaclip_score_histogram")
    return hist


def auc_ap(hist: torch.Tensor) -> Tuple[float, float, int, int]:
    """(AUROC, AP, #positives, #negatives) from a [2, n_bins] count histogram (aaclip_auc_ap)."""
    _need_cuda(hist)
    if hist.dtype == torch.int64:
        if int(hist.max()) >= 2 ** 32:
            raise OverflowError("a histogram bin exceeds 2^32 counts")
        hist = hist.to(torch.int32)      # wraps into the uint32 bit pattern the kernel reads
    hist = hist.contiguous()
    out = torch.empty(4, dtype=torch.float64, device=hist.device)
    with torch.cuda.device(hist.device):
        _lib.check(_lib.load().aaclip_auc_ap(_ptr(hist), hist.shape[1], _ptr(out), _stream()), "aaclip_auc_ap")
    a, p, npos, nneg = out.cpu().tolist()
    return a, p, int(npos), int(nneg)


def row_max(x: torch.Tensor) -> torch.Tensor:
    """[B, ...] fp32 -> [B] maxima (aaclip_row_max)."""
    _need_cuda(x)
    x = x.contiguous().float()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().aaclip_row_max(_ptr(x), x.shape[0], x[0].numel(), _ptr(out), _stream()), "aaclip_row_max")
    return out


def roc_auc_ap(scores: torch.Tensor, labels: torch.Tensor, normalize: bool = False, group=None) -> Tuple[float, float]:
    """AUROC and AP of device scores vs {0,1} labels (what roc_auc_score / average_precision_score return)."""
    mm = reduce_minmax(minmax([scores]), group) if normalize else None
    return auc_ap(reduce_histogram(histogram([scores.reshape(-1)], [labels.reshape(-1)], mm), group))[:2]


def metrics_eval_device(pixel_label: Sequence[torch.Tensor], image_label: torch.Tensor,
                        pixel_preds: Sequence[torch.Tensor], image_preds: torch.Tensor, class_names: str, domain: str,
                        group=None) -> dict:
    """forward_utils.metrics_eval (forward_utils.py:233-308) on device tensors. When torch.distributed is initialised
    with more than one rank (in `group`, or in the default group when group is None) each rank passes ITS shard and
    every rank gets the global result.

    pixel_label / pixel_preds: per-batch lists of [b,(1,)S,S] masks ({0,1}) and fp32 maps; image_label [n] {0,1};
    image_preds [n] (or the fork's [n,2], column 0 is used -- forward_utils.py:269-270).
    """
    preds = [p.reshape(p.shape[0], -1) for p in pixel_preds]
    masks = [m.reshape(m.shape[0], -1) for m in pixel_label]
    mm = reduce_minmax(minmax(preds), group)
    lo, hi = mm.cpu().tolist()
    normalize = hi != 1                                   # `if pixel_preds.max() != 1` (:246)
    hist = reduce_histogram(histogram(preds, masks, mm if normalize else None), group)
    pixel_auc, pixel_ap, _, _ = auc_ap(hist)
    # image level: pmax of the normalised maps (monotone map of the raw maxima), blended with the det score
    pmax = torch.cat([row_max(p) for p in preds])
    if normalize:
        pmax = (pmax - lo) / (hi - lo)
    ip = image_preds.float()
    if ip.dim() == 2 and ip.shape[1] == 2:
        ip = ip[:, 0]
    ip = ip.reshape(-1)
    il = image_label.reshape(-1).to(pmax.device)
    # image-level statistics over ALL ranks' images -- same predicate as the pixel-level reductions above (with the
    # default process group and group=None the pixel metrics used to be global and these shard-local)
    pmax, ip, il = gather_concat(pmax, group), gather_concat(ip.to(pmax.device), group), gather_concat(il.float(), group)
    ip = ip.to(pmax.device)
    if float(ip.max()) != 1:                              # :250-253
        ip = (ip - ip.min()) / (ip.max() - ip.min())
    img = pmax * 0.5 + ip * 0.5 if domain != "Medical" else pmax
    if float(il.max()) != float(il.min()):
        image_auc, image_ap = roc_auc_ap(img, il, normalize=False)
    else:
        image_auc = image_ap = 0
    return {"class name": class_names, "pixel AUC": round(pixel_auc, 4) * 100, "pixel AP": round(pixel_ap, 4) * 100,
            "image AUC": round(image_auc, 4) * 100, "image AP": round(image_ap, 4) * 100}
