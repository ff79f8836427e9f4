"""Pool scan on a real B200: similarity + softmax + argmax against the numpy oracle, and the
leaderboard bit-exact against (a) the golden vectors produced by the reference's own
compute_pseudo_labels and (b) the oracle replay of the kernel's own probabilities."""
import importlib

import numpy as np
import pytest
import torch

from oracle import leaderboard_ref, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod(pkg):
    return importlib.import_module("menghini-neurips23-code_b200.engine")


class _Sim:
    """Engine-free access to the sim kernel (no tower weights needed)."""

    def __init__(self, pkg, eng_mod):
        self.ctx = pkg.Context.get(0)
        self.ptr = importlib.import_module("menghini-neurips23-code_b200._lib").ptr
        self.sp = importlib.import_module("menghini-neurips23-code_b200._lib").stream_ptr

    def __call__(self, F, T, scale, mode=0, probs=True):
        N, C = F.shape[0], T.shape[0]
        pred = torch.empty(N, device="cuda", dtype=torch.int32)
        pp = torch.empty(N, device="cuda", dtype=torch.float32)
        pr = torch.empty(N, C, device="cuda", dtype=torch.float32) if probs else None
        rc = self.ctx.lib.gb_sim_softmax_argmax(self.ctx.h, self.ptr(F), self.ptr(T), scale, N, C, mode,
                                                self.ptr(pred), self.ptr(pp), self.ptr(pr), self.sp())
        self.ctx.check(rc, "sim")
        torch.cuda.synchronize()
        return pred, pp, pr


@pytest.fixture(scope="module")
def sim(pkg, eng_mod):
    return _Sim(pkg, eng_mod)


@pytest.mark.parametrize("N,C", [(1, 1), (5, 10), (128, 16), (129, 45), (1000, 100), (4097, 102),
                                 (300, 128), (777, 7), (20000, 18),
                                 (300, 129), (2000, 200), (1500, 397), (515, 512)])  # class-chunked path
def test_sim_softmax_argmax_matches_oracle(sim, N, C):
    f, t = synth.pool(N, C, peaked=0.05 if N % 2 else 0.3)
    F, T = f.half().cuda(), t.half().cuda()
    pred, pp, probs = sim(F, T, 100.0)
    # oracle on the SAME fp16-rounded inputs, fp32 arithmetic (tolerance: fp32 accumulation order +
    # ex2-based exp: 2e-5 absolute on probabilities)
    # inputs are already unit rows up to fp16 rounding; the kernel does not renormalise
    lg = 100.0 * (F.float().cpu().numpy() @ T.float().cpu().numpy().T)
    e = np.exp(lg - lg.max(1, keepdims=True))
    o_probs = e / e.sum(1, keepdims=True)
    got = probs.cpu().numpy()
    assert np.abs(got - o_probs).max() < 2e-5
    assert np.abs(got.sum(1) - 1).max() < 1e-5
    # argmax: identical wherever the oracle's top-2 margin exceeds the tolerance
    srt = np.sort(o_probs, axis=1)
    clear = (srt[:, -1] - (srt[:, -2] if C > 1 else 0)) > 1e-4
    assert (pred.cpu().numpy()[clear] == o_probs.argmax(1)[clear]).all()
    # pred / p_pred are self-consistent with the probabilities the kernel wrote (bit-exact)
    gp = torch.from_numpy(got)
    assert torch.equal(pred.cpu().long(), gp.argmax(1))
    assert torch.equal(pp.cpu(), gp.max(1).values)
    # without the prob matrix the same pred / p_pred come out
    pred2, pp2, _ = sim(F, T, 100.0, probs=False)
    assert torch.equal(pred2, pred) and torch.equal(pp2, pp)


def test_sim_mode1_argmax_logits(sim):
    f, t = synth.pool(3000, 10, peaked=0.2)
    F, T = f.half().cuda(), t.half().cuda()
    p0, _, _ = sim(F, T, 100.0, mode=0)
    p1, _, _ = sim(F, T, 100.0, mode=1)
    lg = (F.float() @ T.float().t())
    clear = (lg.topk(2, dim=1).values[:, 0] - lg.topk(2, dim=1).values[:, 1]) > 1e-3
    assert torch.equal(p1[clear].long(), lg.argmax(1)[clear])
    assert torch.equal(p0[clear], p1[clear])


def test_sim_rejects_too_many_classes(sim, pkg):
    F = torch.zeros(10, 512, device="cuda", dtype=torch.float16)
    T = torch.zeros(513, 512, device="cuda", dtype=torch.float16)
    with pytest.raises(pkg.GripB200Error):
        sim(F, T, 100.0)


def _golden_cases(golden_dir):
    g = np.load(f"{golden_dir}/leaderboard_cases.npz")
    return g, [str(n) for n in g["names"]]


@pytest.mark.parametrize("prefilter", [False, True])
def test_leaderboard_golden_bit_exact(eng_mod, golden_dir, prefilter):
    g, names = _golden_cases(golden_dir)
    for name in names:
        k = int(g[f"{name}.k"])
        if k == leaderboard_ref.ALL_UNLABELED_K:
            continue
        probs = torch.from_numpy(g[f"{name}.probs"]).cuda().contiguous()
        pred = torch.from_numpy(g[f"{name}.pred"]).to(torch.int32).cuda()
        rank = torch.from_numpy(g[f"{name}.rank"]).to(torch.int32).cuda()
        lb = eng_mod.Leaderboard(probs.shape[1], k, "cuda:0")
        lb.update(probs, pred, rank, prefilter=prefilter)
        idx, lab = lb.result(g[f"{name}.class_ids"].tolist())
        assert idx == g[f"{name}.out_idx"].tolist(), name
        assert lab == g[f"{name}.out_lab"].tolist(), name


@pytest.mark.parametrize("N,C,k,peaked", [(5000, 10, 16, 0.3), (20000, 45, 16, 0.1), (30000, 100, 16, 0.0),
                                          (4000, 7, 64, 0.2), (3000, 5, 600, 0.3), (2000, 3, 1, 0.5),
                                          (10000, 102, 4, 0.05), (6000, 200, 8, 0.1), (3000, 397, 3, 0.1),
                                          # boards too large for shared memory: kept as sets in global memory
                                          (20000, 100, 150, 0.05), (12000, 45, 266, 0.1), (8000, 10, 800, 0.3),
                                          (9000, 128, 65, 0.0), (5000, 2, 2500, 0.2)])
def test_leaderboard_random_matches_oracle(eng_mod, sim, N, C, k, peaked):
    f, t = synth.pool(N, C, peaked=peaked)
    F, T = f.half().cuda(), t.half().cuda()
    pred, pp, probs = sim(F, T, 100.0)
    rank_np = synth.path_ranks(N)
    rank = torch.from_numpy(rank_np).to(torch.int32).cuda()
    want_idx, want_lab = leaderboard_ref.leaderboard(probs.cpu().numpy(), pred.cpu().numpy(), k, rank_np)
    for prefilter in (False, True):
        lb = eng_mod.Leaderboard(C, k, "cuda:0")
        lb.update(probs, pred, rank, prefilter=prefilter)
        idx, lab = lb.result()
        assert idx == want_idx and lab == want_lab, (prefilter, N, C, k)
    # fused scan (filter inside the sim epilogue, candidate rows only)
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    pred2, pp2, _ = lb.scan(F, T, 100.0, rank=rank)
    idx, lab = lb.result()
    assert torch.equal(pred2, pred) and torch.equal(pp2, pp)
    assert idx == want_idx and lab == want_lab, ("fused", N, C, k)


def test_leaderboard_ties_fp16_grid(eng_mod):
    # probabilities on a coarse grid → many exact ties; path rank decides inside sorted()
    rng = np.random.RandomState(3)
    N, C, k = 6000, 8, 12
    lg = rng.choice(np.linspace(0, 3, 7), size=(N, C)).astype(np.float32)
    probs_t = torch.softmax(torch.from_numpy(lg), dim=1)
    pred_t = probs_t.argmax(1)
    rank_np = synth.path_ranks(N, seed=9)
    want = leaderboard_ref.leaderboard(probs_t.numpy(), pred_t.numpy(), k, rank_np)
    for prefilter in (False, True):
        lb = eng_mod.Leaderboard(C, k, "cuda:0")
        lb.update(probs_t.cuda(), pred_t.to(torch.int32).cuda(),
                  torch.from_numpy(rank_np).to(torch.int32).cuda(), prefilter=prefilter)
        assert lb.result() == want


@pytest.mark.parametrize("k", [100, 750, 2000])
def test_leaderboard_ties_large_k(eng_mod, k):
    # the set-mode boards (k > 64): exact ties everywhere, so which entry leaves a full board and the final list
    # order are decided by the path ranks
    rng = np.random.RandomState(11)
    N, C = 6000, 8
    lg = rng.choice(np.linspace(0, 3, 5), size=(N, C)).astype(np.float32)
    probs_t = torch.softmax(torch.from_numpy(lg), dim=1)
    pred_t = probs_t.argmax(1)
    rank_np = synth.path_ranks(N, seed=4)
    want = leaderboard_ref.leaderboard(probs_t.numpy(), pred_t.numpy(), k, rank_np)
    for prefilter in (False, True):
        lb = eng_mod.Leaderboard(C, k, "cuda:0")
        lb.update(probs_t.cuda(), pred_t.to(torch.int32).cuda(),
                  torch.from_numpy(rank_np).to(torch.int32).cuda(), prefilter=prefilter)
        assert lb.result() == want, (k, prefilter)
    # and without a rank array the image index is the tie-break key
    want = leaderboard_ref.leaderboard(probs_t.numpy(), pred_t.numpy(), k, list(range(N)))
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    lb.update(probs_t.cuda(), pred_t.to(torch.int32).cuda(), None, prefilter=True)
    assert lb.result() == want, k


def test_leaderboard_set_mode_without_group_minima(eng_mod, sim, monkeypatch):
    """Boards beyond ≈1.4 M entries do not get the shared-memory group minima and find the entry that leaves by
    scanning the whole board: forced here at a size the oracle can check (random pool and an all-ties pool)."""
    monkeypatch.setenv("GB_LB_NO_GROUPS", "1")
    N, C, k = 9000, 12, 300
    f, t = synth.pool(N, C, peaked=0.1)
    F, T = f.half().cuda(), t.half().cuda()
    rank_np = synth.path_ranks(N)
    rank = torch.from_numpy(rank_np).to(torch.int32).cuda()
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    pred, _, probs = lb.scan(F, T, 100.0, rank=rank, want_probs=True)
    assert lb.result() == leaderboard_ref.leaderboard(probs.cpu().numpy(), pred.cpu().numpy(), k, rank_np)
    rng = np.random.RandomState(12)
    lg = rng.choice(np.linspace(0, 3, 5), size=(5000, 6)).astype(np.float32)
    probs_t = torch.softmax(torch.from_numpy(lg), dim=1)
    pred_t = probs_t.argmax(1)
    rk = synth.path_ranks(5000, seed=3)
    lb = eng_mod.Leaderboard(6, 200, "cuda:0")
    lb.update(probs_t.cuda(), pred_t.to(torch.int32).cuda(), torch.from_numpy(rk).to(torch.int32).cuda(), prefilter=True)
    assert lb.result() == leaderboard_ref.leaderboard(probs_t.numpy(), pred_t.numpy(), 200, rk)


def test_leaderboard_sharded_handoff_large_k(eng_mod, sim):
    """Hand-off of set-mode boards (GRIP's late iterations, k ≈ N/C): every shard's scan call ends by restoring the
    list order, the next shard resumes from it — identical to one scan and to the oracle."""
    N, C, k = 16000, 18, 700
    f, t = synth.pool(N, C, peaked=0.1)
    F, T = f.half().cuda(), t.half().cuda()
    rank_np = synth.path_ranks(N)
    rank = torch.from_numpy(rank_np).to(torch.int32).cuda()
    one = eng_mod.Leaderboard(C, k, "cuda:0")
    pred, _, probs = one.scan(F, T, 100.0, rank=rank, want_probs=True)
    want = one.result()
    assert want == leaderboard_ref.leaderboard(probs.cpu().numpy(), pred.cpu().numpy(), k, rank_np)
    for shards in (2, 5):
        bounds = [N * s // shards for s in range(shards + 1)]
        state = None
        for s in range(shards):
            lb = eng_mod.Leaderboard(C, k, "cuda:0", state=state)
            lb.scan(F[bounds[s]:bounds[s + 1]], T, 100.0, idx0=bounds[s], rank=rank)
            state = lb.state.clone()
        assert lb.result() == want, shards


def test_leaderboard_sharded_handoff_is_identical(eng_mod, sim):
    """Ordered hand-off: shard s continues on the state left by shard s-1 (as rank s would after
    receiving it) — result identical to the single-shard scan for 2, 4 and 8 shards."""
    N, C, k = 24000, 45, 16
    f, t = synth.pool(N, C, peaked=0.1)
    F, T = f.half().cuda(), t.half().cuda()
    rank = torch.from_numpy(synth.path_ranks(N)).to(torch.int32).cuda()
    one = eng_mod.Leaderboard(C, k, "cuda:0")
    one.scan(F, T, 100.0, rank=rank)
    want = one.result()
    for shards in (2, 4, 8, 7):
        bounds = [N * s // shards for s in range(shards + 1)]
        state = None
        for s in range(shards):
            lb = eng_mod.Leaderboard(C, k, "cuda:0", state=state)
            lb.scan(F[bounds[s]:bounds[s + 1]], T, 100.0, idx0=bounds[s], rank=rank)
            state = lb.state.clone()  # what would travel to the next rank
        assert lb.result() == want, shards


def test_leaderboard_rejects_bad_k(eng_mod, pkg):
    with pytest.raises(pkg.GripB200Error):
        eng_mod.Leaderboard(4, 0, "cuda:0")


def test_large_pool_properties(eng_mod, sim):
    """Full-size pool (N = 1,048,576, C = 100, k = 16): size-independent properties instead of the
    (slow) oracle — every board full, entries unique per board, each board's entries are sorted
    descending after its first sort, labels consistent, and identical to the two-step path."""
    N, C, k = 1 << 20, 100, 16
    f, t = synth.pool(N, C, peaked=0.05)
    F, T = f.half().cuda(), t.half().cuda()
    rank = torch.from_numpy(synth.path_ranks(N)).to(torch.int32).cuda()
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    pred, pp, _ = lb.scan(F, T, 100.0, rank=rank)
    idx, ln, p = lb.export(want_p=True)
    assert (ln == k).all()
    idx, p = idx.cpu(), p.cpu()
    for j in range(C):
        assert len(set(idx[j].tolist())) == k
    pred3, pp3, probs = sim(F, T, 100.0)
    assert torch.equal(pred3, pred) and torch.equal(pp3, pp)
    # stored probabilities are the kernel's probabilities of (image, board class)
    want_p = probs[idx.long().cuda().reshape(-1), torch.arange(C, device="cuda").repeat_interleave(k)]
    assert torch.equal(want_p.cpu().reshape(C, k), p)
    lb2 = eng_mod.Leaderboard(C, k, "cuda:0")
    lb2.update(probs, pred, rank, prefilter=True)
    assert lb2.result() == lb.result()
    # the unfiltered replay visits all 2^20 rows strictly in order: pins the pre-filter at full size
    lb3 = eng_mod.Leaderboard(C, k, "cuda:0")
    lb3.update(probs, pred, rank, prefilter=False)
    assert lb3.result() == lb.result()


def test_large_pool_properties_grip_sized_k(eng_mod, sim):
    """GRIP's last iteration at the full pool size (N = 1,048,576, C = 100, k = N/C = 10 485 — boards kept as sets in global
    memory): the oracle would take hours, so the three device routes, which chunk and filter the pool differently, must
    agree entry for entry — fused scan (similarity + pre-filter in growing chunks), replay of the full probability matrix
    with the standalone pre-filter, and the unfiltered replay that visits every row in order — and the boards must have
    the reference's shape: at most k entries, no image twice per board, sorted by (p, path) descending once admitted to."""
    N, C = 1 << 20, 100
    k = N // C
    f, t = synth.pool(N, C, peaked=0.05)
    F, T = f.half().cuda(), t.half().cuda()
    rank = torch.from_numpy(synth.path_ranks(N)).to(torch.int32).cuda()
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    pred, pp, _ = lb.scan(F, T, 100.0, rank=rank)
    idx, ln, p = lb.export(want_p=True)
    assert int(ln.max()) <= k and int(ln.sum()) > N // 2
    pred3, pp3, probs = sim(F, T, 100.0)
    assert torch.equal(pred3, pred)
    for j in (0, 17, 99):
        n = int(ln[j])
        ids = idx[j, :n]
        assert ids.unique().numel() == n
        # stored probabilities are the kernel's probabilities of (image, board class)
        assert torch.equal(probs[ids.long(), j], p[j, :n])
        if n == k:   # full boards that had an admission are sorted; (p, rank) descending
            pj, rj = p[j, :n], rank[ids.long()]
            srt = (pj[:-1] > pj[1:]) | ((pj[:-1] == pj[1:]) & (rj[:-1] > rj[1:]))
            arrival = bool((ids[:-1] < ids[1:]).all())      # never admitted to: still in arrival order
            assert bool(srt.all()) or arrival
    want = lb.result()
    for prefilter in (True, False):
        lb2 = eng_mod.Leaderboard(C, k, "cuda:0")
        lb2.update(probs, pred, rank, prefilter=prefilter)
        assert lb2.result() == want, prefilter


def test_grip_schedule_with_learned_prompts(pkg):
    """GRIP refresh (methods/semi_supervised_learning/pseudo_iterative.py:62-75,113-125 schedule;
    assign_pseudo_labels of textual_fpl.py:195-283): k grows from N/(10·C) towards N/C over the iterations
    and the class decision is argmax(LOGITS) (mode 1).  utils.scan_features against the oracle replay of
    the device probabilities at every k, including the k == 10000000 "label everything" branch."""
    U = importlib.import_module("menghini-neurips23-code_b200.utils")
    ctx = pkg.Context.get(0)

    class _Eng:  # scan_features only needs these three members of an Engine
        device = torch.device("cuda", 0)
        logit_scale_exp = 100.0

        def sim_softmax_argmax(self, F, T, scale=None, mode=0, want_probs=False):
            return _Sim(pkg, None)(F, T, 100.0 if scale is None else scale, mode=mode, probs=want_probs)

    N, C = 9000, 18  # RESICS45 TRZSL: 18 unseen classes
    f, t = synth.pool(N, C, peaked=0.08)
    F, T = f.half().cuda(), t.half().cuda()
    paths = [f"img_{i:06d}.jpg" for i in np.random.RandomState(5).permutation(N)]
    class_ids = [40 - 2 * j for j in range(C)]
    pred, _, probs = _Sim(pkg, None)(F, T, 100.0, mode=1)
    rank = U.path_ranks(paths).numpy()
    num_iter = 10
    for it in (1, 4, 10):
        k = int(it * (N // num_iter) / C)
        idx, lab = U.scan_features(_Eng(), F, T, k, paths, class_ids, mode=1)
        w_idx, w_lab = leaderboard_ref.leaderboard(probs.cpu().numpy(), pred.cpu().numpy(), k, rank, class_ids)
        assert idx == w_idx and lab == w_lab, k
    idx, lab = U.scan_features(_Eng(), F, T, leaderboard_ref.ALL_UNLABELED_K, paths, class_ids, mode=1)
    assert idx == list(range(N)) and lab == [class_ids[j] for j in pred.cpu().tolist()]


def test_fp16_probability_mode_replays_the_references_cuda_arithmetic(pkg, monkeypatch):
    """GRIPB200_PROB_DTYPE=fp16: logits and probabilities in fp16 as the reference's CUDA path holds them
    (utils/clip_pseudolabels.py:59-64 with an fp16 clip_model) — many exact ties, saturation at 1.0 — and the exact
    leaderboard on those numbers; the oracle replays the same fp16 probabilities."""
    U = importlib.import_module("menghini-neurips23-code_b200.utils")

    class _Eng:
        device = torch.device("cuda", 0)
        logit_scale_exp = 100.0

    N, C, k = 12000, 10, 16
    f, t = synth.pool(N, C, peaked=0.3)
    F, T = f.half().cuda(), t.half().cuda()
    paths = [f"img_{i:06d}.jpg" for i in np.random.RandomState(8).permutation(N)]
    rank = U.path_ranks(paths).numpy()
    logits = (torch.tensor(100.0, device="cuda") * F) @ T.t()
    assert logits.dtype == torch.float16
    p16 = logits.softmax(dim=-1)
    ties = int((p16.max(dim=1).values == 1.0).sum())
    assert ties > 50, ties     # the saturation the fp32 path never shows
    for mode in (0, 1):
        pred = (p16 if mode == 0 else logits).argmax(dim=-1)
        want = leaderboard_ref.leaderboard(p16.float().cpu().numpy(), pred.cpu().numpy(), k, rank)
        assert U.scan_features(_Eng(), F, T, k, paths, list(range(C)), mode=mode, probs="fp16") == want
    monkeypatch.setenv("GRIPB200_PROB_DTYPE", "fp16")
    pred0 = p16.argmax(dim=-1)
    want0 = leaderboard_ref.leaderboard(p16.float().cpu().numpy(), pred0.cpu().numpy(), k, rank)
    assert U.scan_features(_Eng(), F, T, k, paths, list(range(C)), mode=0) == want0
    idx, lab = U.scan_features(_Eng(), F, T, leaderboard_ref.ALL_UNLABELED_K, paths, list(range(C)), mode=0)
    assert idx == list(range(N)) and lab == pred0.cpu().tolist()


def test_evaluation_path_matches_the_reference_loop():
    """fused similarity + arg-max over a pool vs the reference's per-batch
    `argmax(logit_scale.exp() * image_features @ text_features.t(), dim=1)` (textual_prompt.py:256-270) evaluated
    in fp32 on the same fp16 unit features; rows whose top-2 logit margin is below fp32 accumulation noise are
    the only ones allowed to differ (none do at these sizes), and the DataFrame has the reference's columns."""
    import importlib

    utils = importlib.import_module("menghini-neurips23-code_b200.utils")
    engine_mod = importlib.import_module("menghini-neurips23-code_b200.engine")
    pkg = importlib.import_module("menghini-neurips23-code_b200")
    f, t = synth.pool(20000, 45, peaked=0.05)
    F, T = f.half().cuda(), t.half().cuda()

    class Eng:  # predict_features only needs the fused pass
        device = torch.device("cuda:0")
        ctx = pkg.Context.get(0)
        lib = ctx.lib
        logit_scale_exp = 100.0
        sim_softmax_argmax = engine_mod.Engine.sim_softmax_argmax

    pred = utils.predict_features(Eng(), F, T).cpu()
    logits = 100.0 * F.float() @ T.float().t()
    want = torch.argmax(logits, dim=1).cpu()
    top2 = torch.topk(logits, 2, dim=1).values.cpu()
    unsure = (top2[:, 0] - top2[:, 1]) < 1e-3
    assert torch.equal(pred.long()[~unsure], want[~unsure])
    assert (pred.long() != want).sum().item() <= unsure.sum().item()
    names = [f"class_{j}" for j in range(45)]
    paths = [f"/data/pool/img_{i % 19990}.png" for i in range(20000)]   # a few repeated file names
    df = utils.predictions_frame(paths, pred, names)
    assert list(df.columns) == ["id", "class"] and df["id"].iloc[0] == "img_0.png"
    assert len(df) == len({(p.split("/")[-1], names[j]) for p, j in zip(paths, pred.tolist())})
