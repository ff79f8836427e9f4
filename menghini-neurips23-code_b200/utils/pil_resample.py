"""Pillow's bicubic `Image.resize` for 8-bit RGB, restated so that it can run on the device bit for bit.

The reference's transform (clip.load's `_transform`, used at data/dataset.py:64-79 and utils/clip_pseudolabels.py:56)
is `Resize(224, BICUBIC) → CenterCrop(224) → …` on PIL images.  Pillow's resampler (src/libImaging/Resample.c, unchanged
in its arithmetic since 3.x; checked here against the installed Pillow by tests/test_pil_resample.py) is integer
arithmetic once the filter coefficients are known:

  * per output coordinate a window [xmin, xmin + n) of source pixels and n double weights of the bicubic kernel
    (a = −0.5, support 2·max(scale, 1) — antialiased when shrinking), normalised to sum 1, then rounded to 22-bit fixed
    point (`PRECISION_BITS = 32 − 8 − 2`);
  * horizontal pass: out = clip8((2²¹ + Σ pixel·k) >> 22) per channel into a uint8 intermediate image;
  * vertical pass: the same on the intermediate.

`coeffs()` computes the windows and fixed-point weights with the same double operations in the same order (Python floats
are C doubles); `resize_crop_np()` is the numpy restatement used by the CPU test and as the checker of the CUDA kernel
(`gb_resize_bicubic_crop_u8`); `clip_geometry()` is torchvision's Resize(224) + CenterCrop(224) size arithmetic.
"""
from __future__ import annotations

import functools
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
_A = -0.5


def _bicubic(x: float) -> float:
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((_A + 2.0) * x - (_A + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * _A
    return 0.0


@functools.lru_cache(maxsize=256)
def coeffs(in_size: int, out_size: int):
    """(bounds int32 [out,2] = (first source index, count), kk int32 [out, ksize]) of Pillow's precompute_coeffs +
    normalize_coeffs_8bpc for a full-extent resize in_size → out_size with the bicubic filter."""
    in0, in1 = 0.0, float(in_size)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, kk


def clip_geometry(w: int, h: int, size: int = 224):
    """torchvision Resize(size) + CenterCrop(size) as clip._preprocess applies them: (new_w, new_h, left, top)."""
    if w <= h:
        nw, nh = size, int(size * h / w)
    else:
        nw, nh = int(size * w / h), size
    left, top = int(round((nw - size) / 2.0)), int(round((nh - size) / 2.0))
    return nw, nh, left, top


def _pass(src: np.ndarray, bounds: np.ndarray, kk: np.ndarray, first: int, count: int) -> np.ndarray:
    """One resampling pass along axis 0 of src [n, m, 3] uint8 for outputs first … first+count−1."""
    out = np.empty((count,) + src.shape[1:], dtype=np.uint8)
    half = 1 << (PRECISION_BITS - 1)
    for o in range(count):
        x0, n = int(bounds[first + o, 0]), int(bounds[first + o, 1])
        acc = np.full(src.shape[1:], half, dtype=np.int64)
        for x in range(n):
            acc += src[x0 + x].astype(np.int64) * int(kk[first + o, x])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_crop_np(img_hwc: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 [H,W,3] → uint8 [3,size,size]: what `clip.preprocess_u8()` returns for the RGB image, computed with the
    integer arithmetic above (horizontal pass first, then vertical, as ImagingResample does)."""
    h, w = img_hwc.shape[:2]
    nw, nh, left, top = clip_geometry(w, h, size)
    x = img_hwc
    if nw != w:   # need_horizontal
        bx, kx = coeffs(w, nw)
        x = _pass(x.transpose(1, 0, 2), bx, kx, left, size).transpose(1, 0, 2)     # [H, size, 3]
    else:
        x = x[:, left:left + size]
    if nh != h:   # need_vertical
        by, ky = coeffs(h, nh)
        x = _pass(x, by, ky, top, size)                                               # [size, size, 3]
    else:
        x = x[top:top + size]
    return np.ascontiguousarray(x.transpose(2, 0, 1))


# ---- decoding in worker PROCESSES ---------------------------------------------------------------------------------------
# PIL releases the GIL inside its decoders, but Image.open / convert / the numpy hand-over are Python: sixteen decoding
# threads top out at ≈3× one thread.  Worker processes scale with the cores.  The staging arenas are anonymous shared
# mappings created BEFORE the workers are forked, so the children inherit them (no names, no resource tracker); they
# never touch CUDA.
_SHARED_ARENAS = []      # mmap objects, index = arena id (inherited by the forked workers)


def _decode_into_arena(args):
    """Worker: decode each path to RGB and copy the pixels into [begin, end) of arena `aid`; returns, per path,
    (offset, H, W) — or None when the region is full (the parent decodes those itself)."""
    aid, begin, end, paths = args
    from PIL import Image
    arena = np.frombuffer(_SHARED_ARENAS[aid], dtype=np.uint8)
    out, off = [], begin
    for pth in paths:
        a = np.asarray(Image.open(pth).convert("RGB"), dtype=np.uint8)
        size = a.size
        if off + size > end:
            out.append(None)
            continue
        arena[off:off + size] = a.reshape(-1)
        out.append((off, a.shape[0], a.shape[1]))
        off += (size + 15) & ~15
    return out


class DeviceResizer:
    """Host side of gb_resize_bicubic_crop_u8.  Two staging arenas (pinned) alternate between chunks; each is split into
    regions, one per decoder: worker PROCESSES (`stage_paths`) or the caller's threads (`put`) copy decoded RGB pixels
    into their region, `flush` uploads the used part of every region into a device mirror of the arena and launches the
    two integer passes once per image size with the images' offsets; coefficient tables are cached per size."""

    def __init__(self, engine, arena_bytes: int = 320 << 20, processes: int = 0):
        import mmap
        import threading

        import torch

        from .._lib import ptr, stream_ptr
        self.torch, self.ptr, self.stream_ptr = torch, ptr, stream_ptr
        self.device = engine.device
        self.lib, self.ctx = engine.ctx.lib, engine.ctx
        self.arena_bytes = int(arena_bytes)
        self.lock = threading.Lock()
        self.tables = {}
        self.uploaded = [None, None]    # event: the device copy of the arena's previous contents has finished
        self.mirror = [None, None]      # device copies of the arenas (offsets are shared)
        self.used = [[], []]            # per slot: [(begin, bytes)] of the regions staged since begin()
        self.off = [0, 0]               # thread mode: bump pointer of the single region
        self.arena_id, self.arenas = [], []
        for _ in range(2):
            m = mmap.mmap(-1, self.arena_bytes)          # anonymous, shared with forked children
            _SHARED_ARENAS.append(m)
            self.arena_id.append(len(_SHARED_ARENAS) - 1)
            t = torch.frombuffer(m, dtype=torch.uint8)
            try:                                          # page-lock it: asynchronous, full-speed host→device copies
                rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), self.arena_bytes, 0)
                self.pinned = int(rc) == 0
            except Exception:
                self.pinned = False
            self.arenas.append(t)
        self.pool, self.processes = None, 0
        if processes and processes > 1:
            try:
                import multiprocessing as mp
                from concurrent.futures import ProcessPoolExecutor
                self.pool = ProcessPoolExecutor(max_workers=int(processes), mp_context=mp.get_context("fork"))
                self.processes = int(processes)
                list(self.pool.map(_decode_into_arena, [(self.arena_id[0], 0, 0, [])] * self.processes))   # fork now
            except Exception:
                self.pool, self.processes = None, 0

    def close(self):
        """Stops the workers and gives the page-locked arenas and their device mirrors back."""
        if self.pool is not None:
            self.pool.shutdown(wait=True, cancel_futures=True)
            self.pool = None
        self.processes = 0
        if self.arenas:
            self.torch.cuda.synchronize(self.device)
            if self.pinned:
                for t in self.arenas:
                    try:
                        self.torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
                    except Exception:
                        pass
            self.arenas, self.mirror = [], [None, None]
            for aid in self.arena_id:            # best effort: the mapping goes once nothing exports its buffer any more
                m, _SHARED_ARENAS[aid] = _SHARED_ARENAS[aid], None
                try:
                    m.close()
                except (BufferError, ValueError, AttributeError):
                    pass

    def _table(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self.tables:
            b, k = coeffs(in_size, out_size)
            self.tables[key] = (self.torch.from_numpy(b.copy()).to(self.device), self.torch.from_numpy(k.copy()).to(self.device),
                                b, k.shape[1])
        return self.tables[key]

    def begin(self, slot):
        """Start staging a chunk into arena `slot` (blocks until the arena's previous upload has finished)."""
        if self.uploaded[slot] is not None:
            self.uploaded[slot].synchronize()
        self.off[slot] = 0
        self.used[slot] = []

    def put(self, slot, array):
        """Thread-safe: copies uint8 [H,W,3] into the arena; returns (offset, H, W), or the array itself when the arena is full."""
        h, w, ch = array.shape
        if ch != 3:
            raise ValueError("DeviceResizer takes RGB images [H,W,3]")
        size = h * w * 3
        with self.lock:
            off = self.off[slot]
            if off + size > self.arena_bytes:
                return array
            self.off[slot] = off + ((size + 15) & ~15)
        self.arenas[slot][off:off + size].view(h, w, 3).copy_(self.torch.from_numpy(array))
        return (off, h, w)

    def stage_paths(self, slot, paths):
        """Decode `paths` in the worker processes straight into arena `slot`; returns what put() would (per path)."""
        from PIL import Image
        n, p = len(paths), self.processes
        region = (self.arena_bytes // p) & ~15
        per = (n + p - 1) // p
        jobs = [(self.arena_id[slot], r * region, (r + 1) * region, paths[r * per:(r + 1) * per]) for r in range(p)]
        puts = []
        for r, res in enumerate(self.pool.map(_decode_into_arena, jobs)):
            top = max([o + ((h * w * 3 + 15) & ~15) for o, h, w in [e for e in res if e is not None]], default=r * region)
            if top > r * region:
                self.used[slot].append((r * region, top - r * region))
            puts.extend(res)
        for i, e in enumerate(puts):         # what did not fit a region: decoded here, uploaded on its own by flush()
            if e is None:
                puts[i] = np.array(Image.open(paths[i]).convert("RGB"), dtype=np.uint8)
        return puts

    def _launch(self, base, entries, out):
        """entries: [(slot index in out, offset, H, W)] all relative to device tensor `base`."""
        torch = self.torch
        groups = {}
        for i, off, h, w in entries:
            groups.setdefault((h, w), []).append((i, off))
        st = self.stream_ptr(self.device)
        for (h, w), items in groups.items():
            nw, nh, left, top = clip_geometry(w, h)
            bx = kx = by = ky = None
            ksx = ksy = 0
            if nw != w:
                bx, kx, _, ksx = self._table(w, nw)
            if nh != h:
                by, ky, by_host, ksy = self._table(h, nh)
                row0 = int(by_host[top, 0])
                rows = int(by_host[top + 223, 0] + by_host[top + 223, 1]) - row0
            else:
                row0, rows = top, 224
            n = len(items)
            meta = torch.tensor([[i for i, _ in items], [o for _, o in items]], dtype=torch.int64)
            index = meta[0].to(torch.int32).to(self.device)
            offs = meta[1].to(self.device)
            tmp = torch.empty(int(self.lib.gb_resize_tmp_bytes(n, rows)), dtype=torch.uint8, device=self.device)
            self.ctx.check(self.lib.gb_resize_bicubic_crop_u8(
                self.ctx.h, self.ptr(base), self.ptr(offs), n, h, w, self.ptr(bx), self.ptr(kx), ksx, left, self.ptr(by),
                self.ptr(ky), ksy, top, row0, rows, self.ptr(index), self.ptr(tmp), self.ptr(out), st),
                "gb_resize_bicubic_crop_u8")

    def flush(self, slot, puts, out):
        """puts[i] = what put() / stage_paths() returned for image i of the chunk; fills out[i] (uint8 [3,224,224], device)."""
        torch = self.torch
        stream = torch.cuda.current_stream(self.device)
        regions = list(self.used[slot])
        if self.off[slot]:
            regions.append((0, self.off[slot]))
        entries = [(i, p[0], p[1], p[2]) for i, p in enumerate(puts) if isinstance(p, tuple)]
        if regions:
            if self.mirror[slot] is None:
                self.mirror[slot] = torch.empty(self.arena_bytes, dtype=torch.uint8, device=self.device)
            for begin, nbytes in regions:
                self.mirror[slot][begin:begin + nbytes].copy_(self.arenas[slot][begin:begin + nbytes], non_blocking=True)
            self.uploaded[slot] = torch.cuda.Event()
            self.uploaded[slot].record(stream)
            self._launch(self.mirror[slot], entries, out)
        for i, p in enumerate(puts):         # images that did not fit the arena: one by one through pageable memory
            if not isinstance(p, tuple):
                base = torch.from_numpy(np.ascontiguousarray(p)).to(self.device)
                self._launch(base, [(i, 0, p.shape[0], p.shape[1])], out)
        return out

    def run(self, arrays, out):
        """Convenience (tests): stage, upload and resize a list of arrays in one call."""
        self.begin(0)
        return self.flush(0, [self.put(0, a) for a in arrays], out)


# One resizer per (device, decoder count) for the whole process: arenas (2 × 320 MB page-locked), their device mirrors
# and the forked decoders are worth keeping, and must not multiply with every model that is loaded.
_RESIZERS = {}


def get_resizer(engine, processes: int = 0) -> DeviceResizer:
    key = (engine.device.index, int(processes) if processes and processes > 1 else 0)
    r = _RESIZERS.get(key)
    if r is None or not r.arenas:
        # the thread-mode resizer serves small pools (and single-decoder callers): a third of the staging is plenty
        r = _RESIZERS[key] = DeviceResizer(engine, processes=key[1], arena_bytes=(320 << 20) if key[1] else (96 << 20))
    return r


def close_resizers():
    for r in list(_RESIZERS.values()):
        r.close()
    _RESIZERS.clear()
