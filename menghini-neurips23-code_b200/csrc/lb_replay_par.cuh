// Parallel exact leaderboard replay — included by sim.cu (inside its anonymous namespace, after LbView,
// LbReplayParams and lb_admit).
//
// Reference semantics: utils/clip_pseudolabels.py:72-101 (and the nine assign_pseudo_labels copies).
//
// Boards evolve independently of one another except for ONE bit per image: "was it rejected by the board of
// its own arg-max class" (only then is it offered to the other boards, :83-101).  So board j is owned by warp
// j % 32 (lane l of warp w looks after board w + 32·l); every warp walks the flagged rows in index order but
// stops only at the rows that can matter to one of ITS boards (p_j > lower bound of board j at the start of
// the round — a full board's minimum never decreases), and the owner of the arg-max board publishes its
// accept / reject decision in a shared-memory byte that the other warps wait for only when (a) one of their
// boards would really take the row and (b) the decision is not already implied by the lower bound
// (p_own ≤ bound ⇒ rejected).  A warp waits only on rows ≤ the one it is at, whose owner never waits at that
// row: no cycle, and all 32 warps are resident (one CTA).  Every board sees exactly the operation sequence of
// the single-warp replay, so the result is bit-identical to it.
#pragma once

constexpr int kLbpThreads = 1024;
constexpr int kLbpWarps = 32;
constexpr int kLbpRound = 4096;     // flagged rows per round (relevance masks: 32 warps × 128 words)
constexpr int kLbpFillRound = 256;  // while some board is not full every row is relevant: refresh the bounds sooner
constexpr int kLbpMaxK = 64;
constexpr int kLbpMaxC = 128;

__device__ __forceinline__ int lbp_block_excl_scan(int v, int32_t* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int t = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    s_warp[lane] = t;
  }
  __syncthreads();
  const int base = warp > 0 ? s_warp[warp - 1] : 0;
  total = s_warp[kLbpWarps - 1];
  __syncthreads();
  return base + x - v;
}

inline size_t lb_replay_par_smem_bytes(int C, int k) {
  return (size_t)3 * C * 4 + (size_t)C * k * 8 + (size_t)C * 4 + (size_t)kLbpWarps * 3 * (k + 1) * 4 +
         (size_t)kLbpRound * 4 + (size_t)kLbpWarps * (kLbpRound / 32) * 4 + (size_t)(kLbpWarps + 4 + 8) * 4 +
         kLbpRound + 16;
}

__global__ void __launch_bounds__(kLbpThreads, 1) lb_replay_par_kernel(const LbReplayParams p) {
  extern __shared__ uint8_t lb_smem[];
  const int C = p.C, k = p.k;
  {
    int32_t* hdr = reinterpret_cast<int32_t*>(p.state);
    if (hdr[0] != C || hdr[1] != k) {
      if (threadIdx.x == 0) hdr[2] = 1;
      return;
    }
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = (C + 31) >> 5;  // boards per warp (≤ 4): lane l < nb of warp w ↔ board w + 32·l
  // ---- shared memory carve-up (every region is a multiple of 4 bytes) ----
  float* s_last = reinterpret_cast<float*>(lb_smem);                 // [C] p of the last list entry
  int32_t* s_cnt = reinterpret_cast<int32_t*>(s_last + C);           // [C]
  float* s_lb = reinterpret_cast<float*>(s_cnt + C);                 // [C] lower bounds of this round
  const LbView g = lb_view(p.state, C, k);
  LbView v = g;
  v.ep = s_lb + C;
  v.ei = reinterpret_cast<int32_t*>(v.ep + (size_t)C * k);
  v.srt = v.ei + (size_t)C * k;
  float* scratch = reinterpret_cast<float*>(v.srt + C);              // [32 warps][3][k+1] one-time sort scratch
  v.sp = scratch + (size_t)warp * 3 * (k + 1);
  v.si = reinterpret_cast<int32_t*>(v.sp + k + 1);
  v.sr = v.si + k + 1;
  int32_t* slots = reinterpret_cast<int32_t*>(scratch + (size_t)kLbpWarps * 3 * (k + 1));  // [kLbpRound]
  uint32_t* rel = reinterpret_cast<uint32_t*>(slots + kLbpRound);    // [32][kLbpRound/32]
  int32_t* s_warp = reinterpret_cast<int32_t*>(rel + kLbpWarps * (kLbpRound / 32));  // [32] scan scratch
  int32_t* s_bc = s_warp + kLbpWarps;                                // [4] n, cursor, boards still filling
  int32_t* s_dbg = s_bc + 4;                                         // [8] diagnostics (→ header words 3..7)
  volatile uint8_t* dec = reinterpret_cast<volatile uint8_t*>(s_dbg + 8);  // [kLbpRound] 0 pending, 1 accepted, 2 rejected
  constexpr int kRelLd = kLbpRound / 32;

  if (tid < 8) s_dbg[tid] = 0;
  const long long t_start = clock64();
  for (int i = tid; i < C * k; i += kLbpThreads) { v.ep[i] = g.ep[i]; v.ei[i] = g.ei[i]; }
  for (int j = tid; j < C; j += kLbpThreads) {
    v.srt[j] = g.srt[j];
    const int c = g.cnt[j];
    s_cnt[j] = c;
    s_last[j] = c > 0 ? g.ep[(size_t)j * k + c - 1] : 0.f;
  }
  __syncthreads();

  const int word_end = (p.row_end + 31) >> 5;
  int cursor = p.row_begin >> 5;  // uniform over the CTA
  while (true) {
    // ---- lower bounds of this round from the live boards; is any board still filling? ----
    if (tid == 0) s_bc[2] = 0;
    __syncthreads();
    for (int j = tid; j < C; j += kLbpThreads) {
      float lbv = -INFINITY;
      if (s_cnt[j] >= k) {
        if (v.srt[j]) {
          lbv = v.ep[(size_t)j * k + k - 1];
        } else {
          lbv = INFINITY;
          for (int e = 0; e < k; ++e) lbv = fminf(lbv, v.ep[(size_t)j * k + e]);
        }
      } else {
        atomicAdd(&s_bc[2], 1);
      }
      s_lb[j] = lbv;
    }
    __syncthreads();
    const int cap = s_bc[2] > 0 ? kLbpFillRound : kLbpRound;
    // ---- (a) the next ≤ cap flagged rows of [row_begin,row_end), in index order ----
    int n = 0;
    while (n < cap && cursor < word_end) {
      const int ww = cursor + tid;
      uint32_t f = 0;
      if (ww < word_end) {
        f = p.flags ? p.flags[ww] : 0xffffffffu;
        const int base = ww << 5;
        if (base < p.row_begin) f &= ~((1u << (p.row_begin - base)) - 1u);   // < 32 by construction
        if (base + 32 > p.row_end) f &= (p.row_end - base) >= 32 ? 0xffffffffu : ((1u << (p.row_end - base)) - 1u);
      }
      const int c = __popc(f);
      int total;
      const int ex = lbp_block_excl_scan(c, s_warp, total);
      const int room = cap - n;
      const bool fits = ex + c <= room;   // true for a prefix of the threads (ex is non-decreasing)
      if (fits) {
        int pos = n + ex;
        while (f) {
          const int b = __ffs(f) - 1;
          f &= f - 1;
          slots[pos++] = (ww << 5) + b;
        }
      }
      if (total <= room) {
        n += total;
        cursor += kLbpThreads;
      } else {
        if (!fits && ex <= room) { s_bc[0] = n + ex; s_bc[1] = ww; }  // the first word that does not fit
        __syncthreads();
        n = s_bc[0];
        cursor = s_bc[1];
        __syncthreads();
        break;
      }
    }
    if (n == 0) break;
    if (tid == 0) { s_dbg[2] += n; s_dbg[5] += 1; }
    // ---- (b) clear the relevance masks and the decisions of this round ----
    const int n_words = (n + 31) >> 5;
    for (int i = tid; i < kLbpWarps * n_words; i += kLbpThreads) rel[(i / n_words) * kRelLd + (i % n_words)] = 0u;
    for (int i = tid; i < n; i += kLbpThreads) dec[i] = 0;
    __syncthreads();
    // ---- (c) relevance: bit s of rel[w] ⇔ row slots[s] beats the bound of some board owned by warp w ----
    for (int s = warp; s < n; s += kLbpWarps) {
      const float* prow = p.rows + (size_t)(slots[s] - p.rows_row0) * C;
      for (int gq = 0; gq < nb; ++gq) {
        const int j = lane + 32 * gq;      // the owner warp of board j is j % 32 == lane
        if (j < C && prow[j] > s_lb[j]) atomicOr(&rel[lane * kRelLd + (s >> 5)], 1u << (s & 31));
      }
    }
    __syncthreads();
    // ---- (d) replay: every warp walks its relevant rows in index order ----
    const int my_board = warp + 32 * lane;            // meaningful for lane < nb
    const bool have_board = lane < nb && my_board < C;
    for (int wi = 0; wi < n_words; ++wi) {
      const uint32_t m = rel[warp * kRelLd + wi];
      if (m == 0) continue;
      // lane e fetches what the warp needs of row slot 32·wi + e (one round of loads per 32 slots)
      int row_e = 0, own_e = 0;
      float pown_e = 0.f, val_e[4] = {0.f, 0.f, 0.f, 0.f};
      if ((m >> lane) & 1u) {
        row_e = slots[wi * 32 + lane];
        own_e = p.pred[row_e];
        const float* prow = p.rows + (size_t)(row_e - p.rows_row0) * C;
        pown_e = prow[own_e];
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          if (gq < nb && warp + 32 * gq < C) val_e[gq] = prow[warp + 32 * gq];
      }
      uint32_t mm = m;
      while (mm) {
        const int e = __ffs(mm) - 1;
        mm &= mm - 1;
        const int s = wi * 32 + e;
        const int own = __shfl_sync(0xffffffffu, own_e, e);
        const float p_own = __shfl_sync(0xffffffffu, pown_e, e);
        const int idx = __shfl_sync(0xffffffffu, row_e, e) + p.idx0;
        const float v0 = __shfl_sync(0xffffffffu, val_e[0], e), v1 = __shfl_sync(0xffffffffu, val_e[1], e);
        const float v2 = __shfl_sync(0xffffffffu, val_e[2], e), v3 = __shfl_sync(0xffffffffu, val_e[3], e);
        const float p_mine = lane == 0 ? v0 : lane == 1 ? v1 : lane == 2 ? v2 : v3;
        const bool known_rej = !(p_own > s_lb[own]);  // at or below the bound: rejected whatever happened since
        bool rejected;
        if (lane == 0) atomicAdd(&s_dbg[0], 1);
        if ((own & 31) == warp) {
          // this warp owns the arg-max board: utils/clip_pseudolabels.py:73-82
          const int c_own = s_cnt[own];
          bool acc = true;
          if (c_own < k) {
            __syncwarp();
            if (lane == 0) {
              v.ep[(size_t)own * k + c_own] = p_own;
              v.ei[(size_t)own * k + c_own] = idx;
              s_cnt[own] = c_own + 1;
              s_last[own] = p_own;
            }
            __syncwarp();
          } else if (s_last[own] < p_own) {
            lb_admit(v, own, p_own, idx, p.rank, s_last, lane);
          } else {
            acc = false;
          }
          if (!known_rej && lane == 0) dec[s] = acc ? 1 : 2;
          rejected = !acc;
        } else {
          // would any of my boards take this row if it is offered?  (live state: exact)
          bool need = false;
          if (have_board) need = s_cnt[my_board] < k || s_last[my_board] < p_mine;
          if (!__any_sync(0xffffffffu, need)) continue;
          if (known_rej) {
            rejected = true;
          } else {
            uint8_t d;
            if (lane == 0) atomicAdd(&s_dbg[1], 1);
            while ((d = dec[s]) == 0) __nanosleep(32);
            rejected = d == 2;
          }
        }
        if (rejected) {  // :83-101 — offered to every other board
          bool need = false;
          if (have_board && my_board != own) {
            const int cj = s_cnt[my_board];
            if (cj < k) {
              v.ep[(size_t)my_board * k + cj] = p_mine;
              v.ei[(size_t)my_board * k + cj] = idx;
              s_cnt[my_board] = cj + 1;
              s_last[my_board] = p_mine;
            } else if (s_last[my_board] < p_mine) {
              need = true;
            }
          }
          uint32_t nm = __ballot_sync(0xffffffffu, need);
          while (nm) {
            const int l = __ffs(nm) - 1;
            nm &= nm - 1;
            const float pj = l == 0 ? v0 : l == 1 ? v1 : l == 2 ? v2 : v3;
            if (lane == 0) atomicAdd(&s_dbg[3], 1);
            lb_admit(v, warp + 32 * l, pj, idx, p.rank, s_last, lane);
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {  // diagnostics, accumulated over the launches of a scan: header words 3..7 =
                   // events walked, decisions waited for, flagged rows, spill admissions, kernel clocks / 1024
    int32_t* hdr = reinterpret_cast<int32_t*>(p.state);
    hdr[3] += s_dbg[0]; hdr[4] += s_dbg[1]; hdr[5] += s_dbg[2]; hdr[6] += s_dbg[3];
    hdr[7] += (int32_t)((clock64() - t_start) >> 10);
  }
  for (int i = tid; i < C * k; i += kLbpThreads) { g.ep[i] = v.ep[i]; g.ei[i] = v.ei[i]; }
  for (int j = tid; j < C; j += kLbpThreads) {
    g.srt[j] = v.srt[j];
    const int c = s_cnt[j];
    g.cnt[j] = c;
    float lbv = -INFINITY;
    if (c >= k) {
      if (v.srt[j]) {
        lbv = v.ep[(size_t)j * k + k - 1];
      } else {
        lbv = INFINITY;
        for (int e = 0; e < k; ++e) lbv = fminf(lbv, v.ep[(size_t)j * k + e]);
      }
    }
    g.lb[j] = lbv;
  }
}
