#!/usr/bin/env python
"""images/sec of the ViT-B/32 prompt-tune + pseudolabel hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    (N > 1: python -m torch.distributed.run --nproc-per-node N … bench.py --gpus N …)

One step = one pass of the hot path over one batch of B synthetic 224x224x3 images per GPU, in the
shape of BASELINE.json configs[1] (CoOp textual prompt, EuroSAT C=10, P=16, ViT-B/32) plus the FPL
pseudolabel assignment (k=16) on the same batch:
  1. frozen image tower on the batch                       (clip_model.encode_image, textual_prompt.py:99-103)
  2. text tower with the learnable prefix, with tape       (TextPrefixModel.forward, :94-97)
  3. cosine logits, cross-entropy, backward to the prefix, SGD update   (:98-135)
  4. similarity + softmax + argmax + per-class leaderboard update with the current prompts
                                                           (assign_pseudo_labels, textual_fpl.py:195-283)
Images are raw uint8 pixels (the resized + centre-cropped crop); ToTensor + Normalize run on the device,
fused into the patch gather, bit-identical to the reference's host transform (tests/test_gpu_towers.py).
Steps 2-3 run through `training.CoOpStep` (gb_text_forward with tape → gb_ce_text_grad → gb_text_backward_prefix →
gb_sgd_step, replayed as one CUDA graph at N = 1): no torch / cuBLAS kernel is left in the timed step.
`value` times the step with the batch already resident in HBM; `e2e` feeds every step from pinned host
memory (H2D copy inside the timed region, double buffered) and reads the loss and predictions back;
`e2e_f32` is the same with host-normalised fp32 tensors, what the reference's DataLoader hands over.
Extra keys, measured outside the timed region: at N = 1 `vpt_step` (configs[2] shape: P=16, C=102, forward with tape +
prompt-only backward through the image tower), `upt_step` (configs[4] shape: 4+4 coupled prompts, C=100, both
towers) and `ref_batch16` (the CoOp step at the reference's own BATCH_SIZE 16, fused and through the plain
drop-in classes); at N > 1 `multi_gpu_parity` (C/G prompts per rank → ONE all-gather of the prototypes → two-phase
sharded pool scan with the ordered leaderboard hand-off, boards compared with a single-GPU scan of the same pool).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"

METRIC = "images/sec ViT-B/32 prompt-tune+pseudolabel"


def flop_vit(P=0):
    """Dense forward FLOPs per image of the ViT-B/32 tower with P prompt rows (SURVEY §8d / BASELINE.md §3)."""
    L = 50 + P
    return 2 * 49 * 3072 * 768 + 12 * (24 * L * 768 ** 2 + 4 * L ** 2 * 768) + 2 * 768 * 512


def flop_vit_executed(P=0):
    """… minus what the CLS-only last block skips (out-proj + MLP of the other L−1 rows): what really runs."""
    return flop_vit(P) - (50 + P - 1) * 18 * 768 ** 2


def flop_text(L):
    return 12 * (24 * L * 512 ** 2 + 4 * L ** 2 * 512) + 2 * 512 ** 2


FLOP_VIT_P0 = flop_vit(0)  # 8.818 G


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="images per GPU per step (0: the largest batch ≤ 2048 whose GEMM tile counts fill whole "
                         "waves of the CTA pairs the image tower runs on, see Engine.wave_aligned_batch)")
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--prefix", type=int, default=16)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--cpu-batch", type=int, default=16, help="reference BATCH_SIZE for the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip vpt_step / upt_step / ref_batch16 / parity")
    ap.add_argument("--workload", default="coop", choices=["coop", "vpt"],
                    help="coop: BASELINE configs[1] (the bench line); vpt: configs[2] shape — visual prompt "
                         "(P=16) train step with backward through the image tower, C=102, extra data point")
    ap.add_argument("--no-overlap", action="store_true",
                    help="run the text chain on the image tower's stream instead of beside it")
    ap.add_argument("--sm-limit", type=int, default=144,
                    help="SMs the persistent image-tower GEMMs may use while the text chain runs beside them")
    return ap.parse_args()


def workload_name(a, batch):
    if getattr(a, "workload", "coop") == "vpt":
        return (f"VPT prompt-tune step (P={a.prefix}, C={a.classes}, ViT-B/32, SGD; forward with tape + "
                f"prompt-only backward through the image tower) + FPL pseudolabel leaderboard (k={a.k}) on "
                f"{batch} synthetic 224x224x3 uint8 images per GPU per step")
    return (f"CoOp prompt-tune step (P={a.prefix}, C={a.classes}, ViT-B/32, SGD) + FPL pseudolabel "
            f"leaderboard (k={a.k}) on {batch} synthetic 224x224x3 uint8 images per GPU per step")


def config_dict(a, B, world):
    """`config` of the JSON line: the same keys and values on both arms (the driver compares them)."""
    vpt = a.workload == "vpt"
    overlap = not a.no_overlap and not vpt
    return {"workload": workload_name(a, B), "weights": "random-init ViT-B/32 (seed 1234)",
            "l2_policy": f"inputs larger than L2 ({B * 150528 / 1e6:.0f} MB uint8 image batch per step + 0.3 GB "
                         f"of weights, GBs of activations)",
            "input": "uint8 pixels, ToTensor + Normalize fused into the patch gather on the device "
                     "(bit-identical to host-normalised fp32 input)",
            "batch_per_gpu": f"{B} (largest <= {1024 if vpt else 2048} that fills whole GEMM waves on the "
                             f"{(a.sm_limit if overlap and a.sm_limit > 0 else 148) // 2} CTA pairs in use)",
            "parallelism": (f"dp{world}: image batch and pool sharded, prefix-grad all-reduce, "
                            f"ordered leaderboard hand-off") if world > 1 else "single GPU",
            "streams": (f"image tower on the main stream (GEMM grids capped at {a.sm_limit} SMs), text "
                        f"chain + pseudolabel scan of the same step on a side stream") if overlap
                       else "single stream",
            "text_positions": "positions after EOT skipped (exact under the causal mask)",
            "reference_arm": f"the same step on the host cores (fp32 torch oracle port of the reference path), each "
                             f"step a bounded sample of {a.cpu_batch} images (the reference's BATCH_SIZE)"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (fp32 torch on the host cores)
# --------------------------------------------------------------------------------------------------
def cpu_arm(a, steps, warmup, budget_s=None):
    from oracle import clip_ref, leaderboard_ref, prompt_ref, synth

    synthetic = importlib.import_module(PKG + ".synthetic")
    torch.set_num_threads(os.cpu_count() or 1)
    model = clip_ref.build_model(synthetic.synthetic_state_dict(1234))
    classes = [" ".join(c.split("_")) for c in synth.class_names(a.classes, seed=1)]
    B = a.cpu_batch
    img = synth.images(B, seed=0)
    labels = torch.arange(B) % a.classes
    prefix = synth.text_prefix(a.prefix)
    boards_probs, boards_pred = [], []

    def step(prefix):
        loss, grad, logits = prompt_ref.coop_step(model, prefix, classes, img, labels)
        prefix = prefix - 1e-4 * grad
        probs = torch.softmax(logits, dim=-1)          # textual_fpl.py:226-228
        boards_probs.append(probs.numpy())
        boards_pred.append(torch.argmax(logits, dim=1).numpy())
        return prefix

    for _ in range(warmup):
        prefix = step(prefix)
    boards_probs.clear(); boards_pred.clear()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        prefix = step(prefix)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    import numpy as np
    pr, pd = np.concatenate(boards_probs), np.concatenate(boards_pred)
    leaderboard_ref.leaderboard(pr, pd, a.k, list(range(len(pd))))
    dt = time.perf_counter() - t0
    return {"value": done * B / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} steps of the same step at the reference BATCH_SIZE={B} "
                      f"(fp32 torch CPU oracle of the reference path, {dt:.1f} s)"}, dt / done * 1e3, done


def run_reference(a, rank, world):
    if rank != 0:
        return
    base, ms, done = cpu_arm(a, a.steps, a.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(a, a.batch, world),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if done != a.steps:
        line["steps_completed"] = done   # the 150 s budget cut the run short
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks sampler (NVML) — runs during the timed region
# --------------------------------------------------------------------------------------------------
class Clocks:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)   # the default timed region is ~0.2 s: sample every 10 ms

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {"hbm_gbs": float(p["hbm_gbs"]),
                    "tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))),
                    "tflops_burst": float(p.get("bf16_tflops", p.get("bf16_tflops_sustained"))),
                    "src": "measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops_sustained; burst = bf16_tflops)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1650.0,
            "src": "fallback (B200_PROFILING.md: 6.65 TB/s copy, 1.4 PFLOP/s sustained / 1.65 burst cuBLAS bf16)"}


def event_ms(fn, steps, warmup=1):
    """Median-free simple timing: `warmup` untimed calls, then `steps` calls between two CUDA events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def make_classes(C, seed=1):
    g = torch.Generator().manual_seed(seed)
    words = ["annual", "crop", "forest", "highway", "industrial", "pasture", "river", "lake", "residential",
             "vegetation", "sea", "road", "land", "buildings"]
    return [" ".join(words[int(torch.randint(0, len(words), (1,), generator=g))]
                     for _ in range(int(torch.randint(1, 5, (1,), generator=g)))) + f" {j}"
            for j in range(C)]


# --------------------------------------------------------------------------------------------------
# extras at N = 1: the other BASELINE configs, each with its own executed-FLOP roofline fraction
# --------------------------------------------------------------------------------------------------
def extras_single_gpu(a, model, dev, pk):
    clip = importlib.import_module(PKG + ".clip")
    models = importlib.import_module(PKG + ".models")
    training = importlib.import_module(PKG + ".training")
    Engine = importlib.import_module(PKG + ".engine").Engine
    eng = model.engine
    eng.ctx.set_sm_limit(0)
    out = {}
    gp = torch.Generator().manual_seed(2)
    gi = torch.Generator().manual_seed(321)

    def frac(tflops):
        return {"tflops": tflops, "frac_of_sustained_peak": tflops / pk["tflops"],
                "frac_of_burst_peak": tflops / pk["tflops_burst"]}

    # ---- configs[2]: VPT, P = 16, C = 102 (Flowers102): forward with tape + prompt-only backward ----
    P, C = 16, 102
    B = Engine.wave_aligned_batch(1024, L=50 + P, sms=148)
    classes = make_classes(C, seed=3)
    cie = models.CustomImageEncoder(model.visual)
    ipm = models.ImagePrefixModel(((768 ** -0.5) * torch.randn(P, 768, generator=gp)).to(dev), cie, device=dev)
    with torch.no_grad():
        tfix = model.encode_text(clip.tokenize([f"a photo of a {c}, a type of flower." for c in classes])).float()
    vstep = training.VPTStep(ipm, tfix, lr=1e-4)
    img = torch.randint(0, 256, (B, 3, 224, 224), generator=gi, dtype=torch.uint8).to(dev)
    lab = torch.randint(0, C, (B,), generator=gi).to(dev)
    l0 = eng.ctx.launches
    ms = event_ms(lambda: vstep.step(img, lab), steps=5, warmup=2)
    per_step = (eng.ctx.launches - l0) // 7
    # forward (dense count, P = 16) + dgrad-only backward ≈ the block GEMMs and attention once more
    f_fwd, f_exec = flop_vit(P), flop_vit_executed(P)
    f_bwd = f_exec - 2 * 49 * 3072 * 768 - 2 * 768 * 512 + 12 * 4 * (50 + P) ** 2 * 768  # attention bwd = 2× its fwd
    out["vpt_step"] = {"config": f"configs[2] shape: VPT P={P}, C={C}, batch {B}, uint8 pixels, SGD; image tower forward "
                                 f"with tape + gb_ce_image_grad + prompt-only backward + gb_sgd_step (training.VPTStep)",
                       "images_per_s": B / (ms * 1e-3), "ms_per_step": ms, "gpu_launches_per_step": per_step,
                       "flop_per_image_executed": f_exec + f_bwd,
                       "roofline_executed": frac(B * (f_exec + f_bwd) / (ms * 1e-3) / 1e12),
                       "roofline_dense_fwd_x2": frac(B * 2 * f_fwd / (ms * 1e-3) / 1e12)}
    del vstep, ipm, img, lab
    torch.cuda.empty_cache()

    # ---- configs[4]: UPT, 4 + 4 coupled prompts, C = 100 (FGVCAircraft): both towers with tape + backward ----
    Pt = Pv = 4
    C = 100
    B = Engine.wave_aligned_batch(1024, L=50 + Pv, sms=148)
    classes = make_classes(C, seed=5)
    cte = models.CustomTextEncoder(model, dev, torch.float32)
    torch.manual_seed(4)
    upt = models.UPTModel((0.02 * torch.randn(1, Pt, 512, generator=gp)).to(dev),
                          ((768 ** -0.5) * torch.randn(1, Pv, 768, generator=gp)).to(dev), None, cie, cte, classes, 128,
                          device=dev, dtype=torch.float32)
    opt = torch.optim.SGD(upt.parameters(), lr=1e-4)
    ustep = training.UPTStep(upt, opt)
    img = torch.randint(0, 256, (B, 3, 224, 224), generator=gi, dtype=torch.uint8).to(dev)
    lab = torch.randint(0, C, (B,), generator=gi).to(dev)
    l0 = eng.ctx.launches
    ms = event_ms(lambda: ustep.step(img, lab), steps=5, warmup=2)
    per_step = (eng.ctx.launches - l0) // 7
    ids = cte._prompt_ids(Pt, classes)
    Lt = int(ids.argmax(dim=-1).max().item()) + 1
    fi = flop_vit_executed(Pv)
    fi_b = fi - 2 * 49 * 3072 * 768 - 2 * 768 * 512 + 12 * 4 * (50 + Pv) ** 2 * 768
    ft = flop_text(Lt)
    ft_b = ft - 2 * 512 ** 2 + 12 * 4 * Lt ** 2 * 512
    tot = B * (fi + fi_b) + C * (ft + ft_b)
    out["upt_step"] = {"config": f"configs[4] shape: UPT {Pt}+{Pv} coupled prompts (128-wide 1-layer coupling transformer in "
                                 f"torch), C={C}, batch {B}, uint8 pixels, SGD; both towers with tape + "
                                 f"gb_ce_image_grad (both gradients) + two prompt-only backward passes "
                                 f"(training.UPTStep); text positions after EOT skipped (Lt={Lt})",
                       "images_per_s": B / (ms * 1e-3), "ms_per_step": ms, "gpu_launches_per_step": per_step,
                       "flop_per_step_executed": tot, "roofline_executed": frac(tot / (ms * 1e-3) / 1e12)}
    del ustep, upt, img, lab
    torch.cuda.empty_cache()

    # ---- N2: the host half of the input pipeline — JPEG files → multi-threaded decode + bicubic resize + crop →
    # pinned uint8 staging → device normalise + image tower (utils.encode_pool) ----
    try:
        import shutil
        import tempfile

        import numpy as np
        from PIL import Image

        utils_b200 = importlib.import_module(PKG + ".utils")
        tmp = tempfile.mkdtemp(prefix="gripb200_jpeg_")
        rng = np.random.default_rng(0)
        n_img = 2048
        base = np.kron(rng.integers(0, 256, size=(12, 16, 3)), np.ones((32, 32, 1))).astype(np.float32)  # 384 x 512
        paths = []
        for i in range(64):     # 64 distinct files, linked 32 times each: the decode cost is what matters
            arr = np.clip(base * rng.uniform(0.6, 1.0) + rng.normal(0, 12, size=base.shape), 0, 255).astype(np.uint8)
            pth = os.path.join(tmp, f"img_{i}.jpg")
            Image.fromarray(arr).save(pth, quality=90)
            paths.append(pth)
        paths = [paths[i % 64] for i in range(n_img)]
        transform = clip._preprocess()
        cores = os.cpu_count() or 1
        os.environ["GRIPB200_POOL_CACHE"] = "0"      # every pass must really decode (the pool repeats 64 files)
        res = {}
        try:
            for dr in (False, True):                 # resize + crop on the host threads | on the device (bit-identical)
                for wk in (1, cores):
                    utils_b200.encode_pool(model, paths[:256], transform, dev, workers=wk, device_resize=dr)   # warm-up: pools, arenas
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    utils_b200.encode_pool(model, paths if wk > 1 else paths[:256], transform, dev, workers=wk, device_resize=dr)
                    torch.cuda.synchronize()
                    res[(dr, wk)] = (n_img if wk > 1 else 256) / (time.perf_counter() - t0)
        finally:
            os.environ.pop("GRIPB200_POOL_CACHE", None)
            shutil.rmtree(tmp, ignore_errors=True)
            importlib.import_module(PKG + ".utils.pil_resample").close_resizers()   # workers, pinned arenas, device mirrors
        out["decode_pipeline"] = {"config": f"{n_img} JPEG files (512x384, quality 90) → utils.encode_pool: decode on {cores} host "
                                            f"threads into pinned staging; Pillow's bicubic resize + centre crop bit for bit ON "
                                            f"THE DEVICE (gb_resize_bicubic_crop_u8), ToTensor + Normalize + image tower on the device",
                                  "images_per_s": res[(True, cores)], "images_per_s_one_thread": res[(True, 1)],
                                  "images_per_s_host_resize": res[(False, cores)],
                                  "images_per_s_host_resize_one_thread": res[(False, 1)], "host_threads": cores,
                                  "note": "the reference decodes AND resizes on one thread at batch 1 (utils/clip_pseudolabels.py:55-57); "
                                          "host JPEG decoding, not the tower, bounds a real pool; images_per_s = forked decoder "
                                          "processes + resize on the device, *_host_resize = decoder threads + PIL resize on the host"}
    except Exception as e:   # PIL without a JPEG codec etc.: an extra, never fatal
        out["decode_pipeline"] = {"unavailable": repr(e)}

    # ---- the reference's own BATCH_SIZE: CoOp, B = 16, C = 10, P = 16 ----
    P, C, B = 16, 10, 16
    classes = make_classes(C, seed=1)
    tpm = models.TextPrefixModel((0.02 * torch.randn(1, P, 512, generator=gp)).to(dev), cte, classes, device=dev)
    fused = training.CoOpStep(tpm, lr=1e-4, graph=True)
    img = torch.randint(0, 256, (B, 3, 224, 224), generator=gi, dtype=torch.uint8).to(dev)
    img32 = clip.normalize_u8(img.cpu()).to(dev)
    lab = torch.randint(0, C, (B,), generator=gi).to(dev)

    def fused_step():
        _, fn, _ = eng.vit_forward(img, None, want_feat=False, want_featn=True)
        fused.step(fn, lab)

    ms_fused = event_ms(fused_step, steps=50, warmup=5)
    with torch.no_grad():
        _, fn_cached, _ = eng.vit_forward(img, None, want_feat=False, want_featn=True)
    ms_cached = event_ms(lambda: fused.step(fn_cached, lab), steps=50, warmup=5)
    tpm2 = models.TextPrefixModel((0.02 * torch.randn(1, P, 512, generator=gp)).to(dev), cte, classes, device=dev)
    opt2 = torch.optim.SGD([tpm2.prefix], lr=1e-4)
    scale = eng.logit_scale_exp

    def dropin_step():   # the reference's loop body, textual_prompt.py:93-135, on the drop-in classes
        tf = tpm2(classes)
        tf = tf / tf.norm(dim=-1, keepdim=True)
        with torch.no_grad():
            imf = model.encode_image(img32)
            imf = imf / imf.norm(dim=-1, keepdim=True)
        loss = torch.nn.functional.cross_entropy(scale * imf @ tf.t(), lab)
        loss.backward()
        opt2.step()
        opt2.zero_grad()

    ms_dropin = event_ms(dropin_step, steps=30, warmup=5)
    ids = cte._prompt_ids(P, classes)
    Lt = int(ids.argmax(dim=-1).max().item()) + 1
    out["ref_batch16"] = {"config": f"CoOp step at the reference BATCH_SIZE: B={B}, C={C}, P={P} (launch-latency bound: "
                                    f"M = {B * 50} image rows, {C * Lt} text rows)",
                          "fused_ms": ms_fused, "fused_images_per_s": B / (ms_fused * 1e-3),
                          "fused_cached_features_ms": ms_cached,
                          "dropin_classes_ms": ms_dropin, "dropin_images_per_s": B / (ms_dropin * 1e-3),
                          "note": "fused = gb_vit_forward + training.CoOpStep (one CUDA graph); cached = CoOpStep alone on "
                                  "cached image features (SURVEY §8f N1); dropin = the reference's loop body on "
                                  "models.TextPrefixModel / clip_model.encode_image with torch loss and optimiser"}
    return out


# --------------------------------------------------------------------------------------------------
# N > 1: the north_star's sharded pool path, with parity against one GPU (outside the timed region)
# --------------------------------------------------------------------------------------------------
def multi_gpu_parity(model, dev, rank, world):
    import torch.distributed as dist

    clip = importlib.import_module(PKG + ".clip")
    engine_mod = importlib.import_module(PKG + ".engine")
    gdist = importlib.import_module(PKG + ".dist")
    eng = model.engine
    eng.ctx.set_sm_limit(0)
    N = 131072
    cases = [(18, 16), (18, N // 18 // 10), (100, 16), (100, N // 100)]   # (C, k): RESICS45 unseen / FGVCAircraft; FPL k and GRIP k
    results = []
    g = torch.Generator(device=dev).manual_seed(5)      # the same pool on every rank
    F = torch.nn.functional.normalize(torch.randn(N, 512, device=dev, generator=g), dim=1).half()
    rank_all = torch.randperm(N, generator=torch.Generator().manual_seed(9)).to(torch.int32).to(dev)
    bounds = gdist.shard_bounds(N, world)
    F_local = F[bounds[rank]:bounds[rank + 1]].contiguous()
    for C, k in cases:
        classes = make_classes(C, seed=10 + C)
        ids = clip.tokenize([f"a photo of a {c}." for c in classes])
        mine = gdist.class_shards(C, world)[rank]
        with torch.no_grad():
            if len(mine):
                _, local, _ = eng.text_forward(ids[mine.start:mine.stop], None, want_feat=False, want_featn=True)
            else:
                local = torch.empty(0, 512, device=dev, dtype=torch.float16)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            protos = gdist.gather_prototypes(local, C)     # ONE all-gather of the [C,512] prototypes
            torch.cuda.synchronize()
            t_gather = time.perf_counter() - t0
            # sharded scan: similarity on every rank at once, exact replay handed rank → rank
            tm = {}
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            board = gdist.sharded_pool_scan(F_local, protos, 100.0, k, N, rank_all,
                                            lambda st: engine_mod.Leaderboard(C, k, dev, state=st), mode=1, timings=tm)
            torch.cuda.synchronize()
            t_sharded = time.perf_counter() - t0
            got = board.result()
            # reference: the whole pool on ONE GPU (every rank computes it: results must agree everywhere)
            one = engine_mod.Leaderboard(C, k, dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            one.scan(F, protos, 100.0, mode=1, idx0=0, rank=rank_all)
            torch.cuda.synchronize()
            t_single = time.perf_counter() - t0
            want = one.result()
            # … and the prototypes against the un-sharded text tower
            _, full, _ = eng.text_forward(ids, None, want_feat=False, want_featn=True)
        same = torch.tensor([int(got == want), int(torch.equal(full, protos))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        times = torch.tensor([t_sharded, tm.get("similarity_ms", 0.0), tm.get("replay_ms", 0.0)], device=dev)
        tmax = times.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = times.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        results.append({"n": N, "c": C, "k": k, "identical": bool(same[0].item()),
                        "prototypes_identical": bool(same[1].item()), "pseudolabels": len(got[0]),
                        "allgather_bytes": int(-(-C // world) * world * 512 * 2), "allgather_ms": t_gather * 1e3,
                        "sharded_scan_ms": float(tmax[0].item()) * 1e3,
                        "similarity_ms_per_rank_max": float(tmax[1].item()),
                        "replay_ms_per_rank_mean": float(tsum[2].item()) / world,
                        "single_gpu_scan_ms": t_single * 1e3})
    del F, F_local
    return {"pool": f"N={N} unit fp16 features (seed 5), prompts through the text tower, mode 1 (arg-max of the logits, "
                    f"assign_pseudo_labels)", "world": world, "cases": results,
            "identical": all(r["identical"] and r["prototypes_identical"] for r in results),
            "note": "wall-clock ms between barriers incl. launch + hand-off latency (max over ranks); similarity = "
                    "phase 1 on every rank at once, replay = phase 2 in rank order (CUDA events)"}


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(a, rank, local_rank, world):
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    clip = importlib.import_module(PKG + ".clip")
    models = importlib.import_module(PKG + ".models")
    engine_mod = importlib.import_module(PKG + ".engine")
    gdist = importlib.import_module(PKG + ".dist")
    training = importlib.import_module(PKG + ".training")
    synthetic = importlib.import_module(PKG + ".synthetic")
    if not torch.cuda.is_available():
        raise pkg.GripB200Error("bench.py needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()

    vpt = a.workload == "vpt"
    overlap = not a.no_overlap and not vpt   # the VPT step is one dependent chain: nothing to run beside it
    B, C, P, k = a.batch, a.classes, a.prefix, a.k
    model, _ = clip.load("ViT-B/32", dev, state_dict=synthetic.synthetic_state_dict(1234))
    eng = model.engine
    ctx = eng.ctx
    classes = make_classes(C, seed=1)
    gp = torch.Generator().manual_seed(2)
    cte = models.CustomTextEncoder(model, dev, torch.float16)
    tpm = models.TextPrefixModel((0.02 * torch.randn(1, P, 512, generator=gp)).to(dev), cte, classes, device=dev)
    coop = training.CoOpStep(tpm, lr=1e-4, world=world, graph=True)   # WARMUP_LR of textual_prompt_config.yml
    scale = eng.logit_scale_exp
    # device-resident arm, 3 roofline steps, e2e arm, e2e_f32 arm: every step scans its own slice of image indices
    total_steps = 3 * (a.warmup + a.steps) + 3
    n_total = total_steps * world * B
    rank_all = torch.randperm(n_total, generator=torch.Generator().manual_seed(7)).to(torch.int32).to(dev)
    board = engine_mod.Leaderboard(C, k, dev)
    state = {"step": 0, "board": board}

    gi = torch.Generator().manual_seed(100 + rank)
    host = [torch.randint(0, 256, (B, 3, 224, 224), generator=gi, dtype=torch.uint8).pin_memory()
            for _ in range(2)]
    host_labels = [(torch.randint(0, C, (B,), generator=gi)).pin_memory() for _ in range(2)]
    dev_img = [h.to(dev) for h in host]
    dev_lab = [h.to(dev) for h in host_labels]
    out_pred = torch.empty(B, dtype=torch.int32).pin_memory()
    out_loss = torch.empty(1, dtype=torch.float32).pin_memory()

    if vpt:
        # visual prompt tuning (methods/semi_supervised_learning/visual_prompt.py:115-145): text features once
        # per epoch from the frozen text tower, learnable rows in the IMAGE tower
        cie = models.CustomImageEncoder(model.visual)
        ipm = models.ImagePrefixModel(((768 ** -0.5) * torch.randn(P, 768, generator=gp)).to(dev), cie, device=dev)
        with torch.no_grad():
            tfix = model.encode_text(clip.tokenize([f"a photo of a {c}" for c in classes])).float()
            protos_fix = (tfix / tfix.norm(dim=-1, keepdim=True)).half()
        vstep = training.VPTStep(ipm, tfix, lr=1e-4, world=world)

    # Two streams: the frozen image tower runs back to back on the main stream; the text chain of a step (text
    # tower with the learnable prefix, loss, prompt-only backward, SGD — one CUDA graph at N = 1 — and the
    # pseudolabel scan; ~230 small, latency-bound kernels) runs beside the NEXT batch's image tower on a side
    # stream.  Nothing is reordered across a true dependency: prefix(i) → text(i) → loss(i) ← image(i).
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream(device=dev)
    mode = {"overlap": overlap}
    if overlap and a.sm_limit > 0:
        ctx.set_sm_limit(a.sm_limit)

    def scan_step(featn, protos, s):
        idx0 = (s * world + rank) * B

        def scan(st):
            b = engine_mod.Leaderboard(C, k, dev, state=st)
            b.scan(featn, protos, scale, mode=1, idx0=idx0, rank=rank_all)
            return b.state

        if world > 1:
            state["board"].state = gdist.ordered_handoff(state["board"].state, scan, ring=True)
        else:
            scan(state["board"].state)

    def step_vpt(img, labels):
        s = state["step"]
        loss, feat, pred = vstep.step(img, labels, want_pred=True)
        featn, _ = ctx.l2norm512(feat)
        state["pred"] = pred
        scan_step(featn, protos_fix, s)
        state["step"] = s + 1
        return loss

    def step(img, labels):
        if vpt:
            return step_vpt(img, labels)
        s = state["step"]
        overlap = mode["overlap"]
        side = side_stream if overlap else main_stream
        with torch.no_grad():
            _, featn, _ = eng.vit_forward(img, None, want_feat=False, want_featn=True)
        if overlap:
            ev = torch.cuda.Event()
            ev.record(main_stream)
            featn.record_stream(side)
        with torch.cuda.stream(side):
            if overlap:
                side.wait_event(ev)
            loss, pred = coop.step(featn, labels, want_pred=True)
            protos, _ = ctx.l2norm512(coop._text)   # unit fp16 prompts of THIS step's forward
            state["pred"] = pred
            scan_step(featn, protos, s)
        state["step"] = s + 1
        return loss

    def join():
        main_stream.wait_stream(side_stream)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record()
        for i in range(steps):
            fn(i)
        join()  # the side stream's work of every step belongs to the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launches - l0

    # ---- device-resident arm -------------------------------------------------------------------
    for i in range(a.warmup):
        step(dev_img[i % 2], dev_lab[i % 2])
    join()
    torch.cuda.synchronize()
    graph_launches = 0
    if not vpt and coop._graph is not None:
        # a replayed CUDA graph launches its captured kernels without passing through the C ABI's counter
        l0 = ctx.launches
        coop._chain()
        graph_launches = ctx.launches - l0
        torch.cuda.synchronize()
    with Clocks(local_rank) as clk:
        ms, launches = timed(lambda i: step(dev_img[i % 2], dev_lab[i % 2]), a.steps)
    launches += graph_launches * a.steps
    value = a.steps * B * world / (ms / 1e3)

    # ---- roofline: per-launch CUDA-event timing (on the launching stream) of every tcgen05 GEMM and sim
    # launch in 3 more steps of the same loop; the dominant kernel launch = the GEMM shape with the
    # largest total time
    # (issued on ONE stream with the SM cap lifted, so a launch's event pair times that kernel alone; the per-launch
    # event pairs also break the programmatic-dependent-launch overlap between consecutive GEMMs, which is why the
    # shares below can add up to slightly more than one step)
    join()
    mode["overlap"] = False
    ctx.set_sm_limit(0)
    use_graph = coop.use_graph
    coop.use_graph, coop._graph = False, None   # eager chain: every launch is visible to the profiler hooks
    ctx.profile_begin()
    for i in range(3):
        step(dev_img[i % 2], dev_lab[i % 2])
    launches_rec = ctx.profile_launches()
    (g_n, g_ms, g_flop), (s_n, s_ms, s_bytes) = ctx.profile_end()
    coop.use_graph = use_graph
    mode["overlap"] = overlap
    if overlap and a.sm_limit > 0:
        ctx.set_sm_limit(a.sm_limit)
    pk = peaks()
    by_shape = {}
    for kind, gm, gn, gk, lms, work in launches_rec:
        if kind == 0:
            e = by_shape.setdefault((gm, gn, gk), [0, 0.0, 0.0])
            e[0] += 1; e[1] += lms; e[2] += work
    (dm, dn, dk), (d_cnt, d_ms, d_flop) = max(by_shape.items(), key=lambda kv: kv[1][1])
    achieved = d_flop / (d_ms * 1e-3) / 1e12
    all_gemm = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"gemm_{dm}x{dn}x{dk}")
    step_s = ms / a.steps * 1e-3
    fP = P if vpt else 0
    roofline = {"bound": "tensor", "kernel": f"gemm_f16_tcgen05_2cta_kernel M={dm} N={dn} K={dk}",
                "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                "frac_of_burst_peak": achieved / pk["tflops_burst"], "peak_burst": pk["tflops_burst"],
                "traffic": traffic, "peak_source": pk["src"],
                "flop_per_launch": 2.0 * dm * dn * dk, "avg_launch_ms": d_ms / d_cnt, "launches_timed": d_cnt,
                "share_of_step": d_ms / 3 / (ms / a.steps),
                "per_shape": {f"{m_}x{n_}x{k_}": {"launches": v[0], "tflops": v[2] / (v[1] * 1e-3) / 1e12,
                                                  "frac": v[2] / (v[1] * 1e-3) / 1e12 / pk["tflops"],
                                                  "share_of_step": v[1] / 3 / (ms / a.steps)}
                              for (m_, n_, k_), v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:6]},
                "all_gemm_launches": {"achieved": all_gemm, "frac": all_gemm / pk["tflops"], "launches": g_n,
                                      "share_of_step": g_ms / 3 / (ms / a.steps),
                                      "note": "per-launch event pairs serialise the launches (no PDL overlap): the shares "
                                              "of a profiled step can add up to more than the un-profiled step time"},
                "step_vit_fwd_frac_of_peak": B * flop_vit(fP) / step_s / 1e12 / pk["tflops"],
                "step_vit_fwd_frac_of_peak_executed": B * flop_vit_executed(fP) / step_s / 1e12 / pk["tflops"],
                "step_vit_fwd_frac_of_burst_peak_executed": B * flop_vit_executed(fP) / step_s / 1e12 / pk["tflops_burst"],
                "step_vit_fwd_flop_note": "dense = the reference count, 8.818 GFLOP per image (SURVEY §8d), over the whole "
                                          "step time; executed = minus the 5.9 % the CLS-only last block skips "
                                          "(bit-identical features); both count the image tower's forward only"}

    # pool-scale sim kernel (HBM bound): N = 2^20 rows against C=100 prototypes, timed alone
    roofline_sim = None
    if rank == 0:
        Np, Cp = 1 << 20, 100
        gpool = torch.Generator(device=dev).manual_seed(5)   # the pool of tools/gpu_scan_*.py: the replay's time depends on the data
        F = torch.nn.functional.normalize(torch.randn(Np, 512, device=dev, generator=gpool), dim=1).half()
        T = torch.nn.functional.normalize(torch.randn(Cp, 512, device=dev, generator=gpool), dim=1).half()
        for _ in range(20):  # long enough for the clocks to settle after the GEMM-heavy steps
            eng.sim_softmax_argmax(F, T, 100.0)
        torch.cuda.synchronize()
        ctx.profile_begin()
        for _ in range(20):
            eng.sim_softmax_argmax(F, T, 100.0)
        (_, _, _), (n1, ms1, by1) = ctx.profile_end()
        gbs = by1 / (ms1 * 1e-3) / 1e9
        # the whole fused pool scan (chunked similarity + pre-filter + exact leaderboard replay), k = 16
        rk = torch.randperm(Np, generator=torch.Generator().manual_seed(9)).to(torch.int32).to(dev)
        scan_ms = []
        for _ in range(3):
            lbp = engine_mod.Leaderboard(Cp, 16, dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            lbp.scan(F, T, 100.0, rank=rk)
            t1.record()
            torch.cuda.synchronize()
            scan_ms.append(t0.elapsed_time(t1))
        scan_ms = sorted(scan_ms)[1]
        roofline_sim = {"bound": "hbm", "kernel": "sim_softmax_argmax_kernel",
                        "workload": f"pool N={Np}, C={Cp}, fp16 features (1 GiB > L2)", "achieved": gbs,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                        "traffic": (json.load(open(tpath)).get("sim_1048576x100") if os.path.exists(tpath) else None),
                        "launches_timed": n1, "images_per_s": Np * n1 / (ms1 * 1e-3),
                        "full_scan_k16": {"ms": scan_ms, "images_per_s": Np / (scan_ms * 1e-3),
                                          "hbm_frac": Np * 1032.0 / (scan_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                          "note": "gb_pseudolabel_scan: every chunk's similarity kernel + "
                                                  "pre-filter + exact leaderboard replay (boards replayed in "
                                                  "parallel by their owner warps)"}}
        del F, T, rk

    # ---- end-to-end arm: pinned host → device every step, loss + predictions read back ---------
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    side = side_stream if overlap else main_stream

    NBUF = 3   # device staging buffers: a copy never has to wait for the side-stream tail of the step before last

    def run_e2e(host_img, dev_bufs):
        # Three device buffers (two pinned host batches feed them alternately).  With two, the copy of batch i+1 had to wait
        # until the side stream had finished step i-1 — whose text chain and scan run BESIDE image tower i on the few SMs
        # the tower leaves free and end late in it — so the 296 MB copy finished just after tower i and tower i+1 waited
        # for it (8 % of the step); the copies themselves do not slow the tower (measured with the copy switched off).
        while len(dev_bufs) < NBUF:
            dev_bufs.append(torch.empty_like(dev_bufs[0]))
        labs = [torch.empty_like(dev_lab[0]) for _ in range(NBUF)]
        copied = [torch.cuda.Event() for _ in range(NBUF)]
        consumed = [torch.cuda.Event() for _ in range(NBUF)]
        consumed_side = [torch.cuda.Event() for _ in range(NBUF)]  # labels are read by the side stream

        def prefetch(i):
            b = i % NBUF
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])
                copy_stream.wait_event(consumed_side[b])
                if os.environ.get("GB_E2E_DIAG_NOCOPY") != "1":   # diagnosis only: what the copies themselves cost
                    dev_bufs[b].copy_(host_img[i % 2], non_blocking=True)
                labs[b].copy_(host_labels[i % 2], non_blocking=True)
                copied[b].record(copy_stream)

        def e2e_step(i):
            b = i % NBUF
            prefetch(i + 1)               # next batch streams in while this one computes
            main.wait_event(copied[b])
            loss = step(dev_bufs[b], labs[b])
            consumed[b].record(main)
            with torch.cuda.stream(side):
                consumed_side[b].record(side)
                out_loss.copy_(loss.detach().reshape(1), non_blocking=True)
                out_pred.copy_(state["pred"], non_blocking=True)
            # the caller reads loss / predictions every step: with the text chain overlapped the values read
            # here are those of the previous step (its side-stream work is what we wait for)
            if overlap:
                if state.get("prev_done") is not None and os.environ.get("GB_E2E_DIAG_NOSYNC") != "1":
                    state["prev_done"].synchronize()
                state["prev_done"] = torch.cuda.Event()
                state["prev_done"].record(side)
            else:
                main.synchronize()

        join()
        torch.cuda.synchronize()
        for b in range(NBUF):
            consumed[b].record(main)
            consumed_side[b].record(main)
        prefetch(0)
        for i in range(a.warmup):
            e2e_step(i)
        ms_, _ = timed(lambda i: e2e_step(i + a.warmup), a.steps)
        torch.cuda.synchronize()
        return ms_

    ms_e2e = run_e2e(host, dev_img)
    e2e_value = a.steps * B * world / (ms_e2e / 1e3)
    h2d = B * 3 * 224 * 224 + B * 8
    d2h = 4 + B * 4
    # the same with the fp32 tensors the reference's DataLoader yields (4x the bytes over PCIe)
    host_f32 = [clip.normalize_u8(h).pin_memory() for h in host]
    dev_f32 = [torch.empty(B, 3, 224, 224, device=dev) for _ in range(NBUF)]
    ms_e2e_f32 = run_e2e(host_f32, dev_f32)
    e2e_f32 = {"value": a.steps * B * world / (ms_e2e_f32 / 1e3), "unit": "images/s",
               "ms_per_step": ms_e2e_f32 / a.steps, "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 8,
               "d2h_bytes_per_step": d2h,
               "note": "host-normalised fp32 tensors (the reference DataLoader's output) instead of uint8 pixels"}
    del host_f32, dev_f32
    host.clear(); dev_img.clear()
    join()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    extras = {}
    if not a.no_extras:
        if world == 1:
            extras = extras_single_gpu(a, model, dev, pk)
        else:
            extras = {"multi_gpu_parity": multi_gpu_parity(model, dev, rank, world)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
                "data": "synthetic", "config": config_dict(a, B, world),
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / a.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "e2e_f32": e2e_f32,
                "roofline": roofline, "roofline_sim": roofline_sim}
        line.update(extras)
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(a, 1000, 1, budget_s=15.0)[0]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.batch <= 0:  # both arms name the same workload
        vpt = a.workload == "vpt"
        capped = not a.no_overlap and not vpt and a.sm_limit > 0
        a.batch = importlib.import_module(PKG + ".engine").Engine.wave_aligned_batch(
            1024 if vpt else 2048, L=50 + (a.prefix if vpt else 0), sms=a.sm_limit if capped else 148)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus:
        if a.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run "
                             f"--nproc-per-node {a.gpus}")
    run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
