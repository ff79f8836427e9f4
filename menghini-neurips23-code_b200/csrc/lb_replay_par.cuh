// Parallel exact leaderboard replay — included by sim.cu (inside its anonymous namespace, after LbView,
// LbReplayParams and lb_admit).
//
// Reference semantics: utils/clip_pseudolabels.py:72-101 (and the nine assign_pseudo_labels copies).
//
// Boards evolve independently of one another except for ONE bit per image: "was it rejected by the board of
// its own arg-max class" (only then is it offered to the other boards, :83-101).  So board j is owned by warp
// j % 32 (lane l of warp w looks after board w + 32·l); every warp goes through the flagged rows in index order,
// 32 at a time, but looks only at the rows that can matter to one of ITS boards (p_j > lower bound of board j at
// the start of the round — a full board's minimum never decreases).  The owner of a row's arg-max board decides it
// from the live board and publishes accept / reject in a shared-memory byte; the other warps read the bytes of the
// rows between two rows of their own all at once (lane e polls slot e; a verdict implied by the lower bound —
// p_own ≤ bound ⇒ rejected — is not waited for), screen the rejected ones against the live thresholds of their
// boards and walk only those a board would take.  A warp waits only on rows before the own row it is about to
// decide, whose owners never wait at or beyond their row: no cycle, and all 32 warps are resident (one CTA).
// Every board sees exactly the operation sequence of the single-warp replay, so the result is bit-identical to it.
//
// kSet = true is the same walk for boards too large for shared memory (GRIP grows k to N/C,
// methods/semi_supervised_learning/pseudo_iterative.py:62-75): the boards stay in the caller's state (global
// memory / L2) and a FULL board is kept as an unordered SET.  That is exact because the reference's
// `sorted(board + [new], reverse=True)[:k]` (:78-82) always drops the smallest (p, path) of the k+1 items and the
// new item is never that one (it was admitted because own[-1].p < p), on the first admission of a board as well as
// on every later one — so an admission is "replace the set's minimum", the only thing that distinguishes the
// never-sorted state from the sorted one is the threshold (p of the k-th ARRIVAL, which never moves in the set
// layout, vs. the set's minimum p), and the list ORDER the reference would show is recovered by one sort at the end
// of the scan call (lb_set_sort_kernel): no O(k²) one-time sort and no O(k) shifting inside the sequential walk.
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

// decision bytes | slots | relevance masks | scan scratch, broadcast words, diagnostics | phase clocks, 16-byte aligned
constexpr size_t kLbpFixedBytes =
    ((size_t)kLbpRound + (size_t)kLbpRound * 4 + (size_t)kLbpWarps * (kLbpRound / 32) * 4 + (size_t)(kLbpWarps + 4 + 8) * 4 + 32 + 15) &
    ~size_t(15);
inline size_t lb_replay_par_smem_bytes(int C, int k, bool set_mode = false) {
  const size_t boards = set_mode ? 0 : (size_t)C * k * 8 + (size_t)C * 4 + (size_t)kLbpWarps * 3 * (k + 1) * 4;
  return kLbpFixedBytes + (size_t)5 * C * 4 + boards + 16;
}
// kSet: minima of every group of 32 consecutive board slots, [C][ceil(k/32)] floats, when they fit beside the rest
// (C·k ≲ 1.4 M entries): an admission then reads one or two 128-byte groups instead of the whole board.
constexpr size_t kLbpSmemMax = 200 * 1024;
inline int lb_set_groups(int C, int k) {
  const int G = (k + 31) / 32;
  return lb_replay_par_smem_bytes(C, k, true) + (size_t)C * G * 4 <= kLbpSmemMax ? G : 0;
}

// decision bytes: volatile shared-space accesses through a 32-bit shared address
__device__ __forceinline__ uint32_t lbp_lds_u8(uint32_t a) {
  uint32_t x;
  asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(x) : "r"(a) : "memory");
  return x;
}
__device__ __forceinline__ uint32_t lbp_lds_u32(uint32_t a) {   // boards' counts / thresholds: written by the owner warp only
  uint32_t x;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a) : "memory");
  return x;
}
__device__ __forceinline__ void lbp_sts_u8(uint32_t a, uint32_t x) {
  asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}

__device__ __forceinline__ float lbp_warp_min(float m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  return m;
}

// Set-mode admission with group minima gm[0..G) of board j in shared memory: the smallest p of the board is the
// smallest group minimum; the entry that leaves is, among the entries that carry it (one group unless p ties
// exactly), the one with the lowest path rank.
__device__ void lb_admit_set_gm(const LbView& v, int j, float pj, int idx, const int32_t* rank, float* s_last,
                                float* gm, int G, int lane) {
  const int k = v.k;
  float* ep = v.ep + (size_t)j * k;
  int32_t* ei = v.ei + (size_t)j * k;
  __syncwarp();
  float m = INFINITY;
  for (int g = lane; g < G; g += 32) m = fminf(m, gm[g]);
  m = lbp_warp_min(m);
  int best_r = 0x7fffffff, best_pos = -1;
  for (int g0 = 0; g0 < G; g0 += 32) {
    const int g = g0 + lane;
    uint32_t bal = __ballot_sync(0xffffffffu, g < G && gm[g] == m);
    while (bal) {
      const int t = __ffs(bal) - 1;
      bal &= bal - 1;
      const int e = (g0 + t) * 32 + lane;
      const float pe = e < k ? ep[e] : INFINITY;
      int r = 0x7fffffff, pos = e;
      if (pe == m) {
        const int ie = ei[e];
        r = rank ? rank[ie] : ie;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int orr = __shfl_xor_sync(0xffffffffu, r, o);
        const int opos = __shfl_xor_sync(0xffffffffu, pos, o);
        if (orr < r || (orr == r && opos < pos)) { r = orr; pos = opos; }
      }
      if (best_pos < 0 || r < best_r) { best_r = r; best_pos = pos; }
    }
  }
  const int gb = best_pos >> 5;
  const int e = gb * 32 + lane;
  float pe = e < k ? ep[e] : INFINITY;
  if (e == best_pos) pe = pj;
  const float gnew = lbp_warp_min(pe);
  if (lane == 0) {
    ep[best_pos] = pj;
    ei[best_pos] = idx;
    gm[gb] = gnew;
    v.srt[j] = 1;
  }
  __syncwarp();
  m = INFINITY;
  for (int g = lane; g < G; g += 32) m = fminf(m, gm[g]);
  m = lbp_warp_min(m);
  if (lane == 0) s_last[j] = m;
  __syncwarp();
}

// Admission of (pj, idx) into FULL board j kept as a set (kSet): the entry that is smallest by (p, path rank) leaves,
// the new one takes its slot; the board's threshold becomes the smallest p that is left.  Warp-cooperative, one
// pass over the k entries (coalesced; ranks are looked at only on exact ties of p).
__device__ void lb_admit_set(const LbView& v, int j, float pj, int idx, const int32_t* rank, float* s_last,
                             int lane) {
  const int k = v.k;
  float* ep = v.ep + (size_t)j * k;
  int32_t* ei = v.ei + (size_t)j * k;
  __syncwarp();
  float m1p = INFINITY, m2p = INFINITY;   // smallest (p, rank) seen by this lane; smallest p among its other entries
  int m1r = 0x7fffffff, m1pos = -1;
  for (int e = lane; e < k; e += 32) {
    const float pe = ep[e];
    bool less = pe < m1p;
    int re = 0;
    if (pe == m1p) {
      const int ie = ei[e];
      re = rank ? rank[ie] : ie;
      less = re < m1r;
    } else if (less) {
      const int ie = ei[e];
      re = rank ? rank[ie] : ie;
    }
    if (less) { m2p = m1p; m1p = pe; m1r = re; m1pos = e; }
    else m2p = fminf(m2p, pe);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float op = __shfl_xor_sync(0xffffffffu, m1p, o);
    const int orr = __shfl_xor_sync(0xffffffffu, m1r, o);
    const int opos = __shfl_xor_sync(0xffffffffu, m1pos, o);
    const float o2 = __shfl_xor_sync(0xffffffffu, m2p, o);
    const bool other_wins = op < m1p || (op == m1p && (orr < m1r || (orr == m1r && opos >= 0 && (m1pos < 0 || opos < m1pos))));
    if (other_wins) { m2p = fminf(fminf(m2p, o2), m1p); m1p = op; m1r = orr; m1pos = opos; }
    else m2p = fminf(fminf(m2p, o2), op);
  }
  if (lane == 0) {
    ep[m1pos] = pj;
    ei[m1pos] = idx;
    v.srt[j] = 1;
    s_last[j] = fminf(m2p, pj);
  }
  __syncwarp();
}

template <bool kSet>
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
  // ---- shared memory carve-up ----
  // The regions the sequential walk touches on its critical path sit at FIXED offsets (decision bytes first): their
  // addresses are constants instead of arithmetic on C and k that the compiler — capped at 64 registers by the 1024
  // threads — re-did, with two special-register reads for the generic-to-shared conversion, in every iteration of the
  // poll loop.
  volatile uint8_t* dec = reinterpret_cast<volatile uint8_t*>(lb_smem);     // [kLbpRound] 0 pending, 1 accepted, 2 rejected
  // (opaque moves: the compiler otherwise re-derives these addresses — a special-register read each — wherever it
  // needs them instead of keeping them in a register)
  uint32_t dec_s, last_s;
  asm volatile("mov.u32 %0, %1;" : "=r"(dec_s) : "r"(smem_u32(lb_smem)));
  asm volatile("mov.u32 %0, %1;" : "=r"(last_s) : "r"(smem_u32(lb_smem + kLbpFixedBytes)));   // s_last; s_cnt follows
  int32_t* slots = reinterpret_cast<int32_t*>(lb_smem + kLbpRound);         // [kLbpRound]
  uint32_t* rel = reinterpret_cast<uint32_t*>(slots + kLbpRound);           // [32][kLbpRound/32]
  int32_t* s_warp = reinterpret_cast<int32_t*>(rel + kLbpWarps * (kLbpRound / 32));  // [32] scan scratch
  int32_t* s_bc = s_warp + kLbpWarps;                                // [4] n, cursor, boards still filling, … nearly full
  int32_t* s_dbg = s_bc + 4;                                         // [8] diagnostics (→ header words 3..7)
  long long* s_clk = reinterpret_cast<long long*>(s_dbg + 8);        // [4] diag 2: last stamp, clocks of phases a / c / d
  float* s_last = reinterpret_cast<float*>(lb_smem + kLbpFixedBytes);  // [C] acceptance threshold: p of the last list entry
  int32_t* s_cnt = reinterpret_cast<int32_t*>(s_last + C);           // [C]
  float* s_lb = reinterpret_cast<float*>(s_cnt + C);                 // [C] lower bounds of this round
  float* s_min = s_lb + C;                                           // [C] kSet: running minimum of a never-sorted board
  int32_t* s_ka = reinterpret_cast<int32_t*>(s_min + C);             // [C] own arrivals of this round are sure to be appended
  const LbView g = lb_view(p.state, C, k);
  LbView v = g;
  float* carve = reinterpret_cast<float*>(s_ka + C);
  if constexpr (!kSet) {
    v.ep = carve;
    v.ei = reinterpret_cast<int32_t*>(v.ep + (size_t)C * k);
    v.srt = v.ei + (size_t)C * k;
    float* scratch = reinterpret_cast<float*>(v.srt + C);            // [32 warps][3][k+1] one-time sort scratch
    v.sp = scratch + (size_t)warp * 3 * (k + 1);
    v.si = reinterpret_cast<int32_t*>(v.sp + k + 1);
    v.sr = v.si + k + 1;
    carve = scratch + (size_t)kLbpWarps * 3 * (k + 1);
  }
  constexpr int kRelLd = kLbpRound / 32;
  const int G = kSet ? p.set_groups : 0;    // group minima of the set-mode boards ([C][G]; 0: not kept)
  float* s_gm = reinterpret_cast<float*>(
      lb_smem + ((reinterpret_cast<uint8_t*>(carve) - lb_smem + 15) & ~size_t(15)));

  if (tid < 8) s_dbg[tid] = 0;
  const long long t_start = clock64();
  if constexpr (kSet) {
    // boards stay in global memory; a board's threshold is the k-th arrival's p until its first admission and the
    // minimum of the set afterwards (entries of an already sorted board may arrive in any order)
    for (int j = warp; j < C; j += kLbpWarps) {
      const int c = g.cnt[j];
      float m = INFINITY;
      if (G > 0) {
#pragma unroll 4
        for (int gq = 0; gq < G; ++gq) {
          const int e = gq * 32 + lane;
          const float gmin = lbp_warp_min(e < c ? g.ep[(size_t)j * k + e] : INFINITY);
          if (lane == 0) s_gm[(size_t)j * G + gq] = gmin;
          m = fminf(m, gmin);
        }
      } else {
        for (int e = lane; e < c; e += 32) m = fminf(m, g.ep[(size_t)j * k + e]);
        m = lbp_warp_min(m);
      }
      if (lane == 0) {
        s_cnt[j] = c;
        s_min[j] = m;
        s_last[j] = (c >= k && g.srt[j]) ? m : (c > 0 ? g.ep[(size_t)j * k + c - 1] : 0.f);
      }
    }
  } else {
    for (int i = tid; i < C * k; i += kLbpThreads) { v.ep[i] = g.ep[i]; v.ei[i] = g.ei[i]; }
    for (int j = tid; j < C; j += kLbpThreads) {
      v.srt[j] = g.srt[j];
      const int c = g.cnt[j];
      s_cnt[j] = c;
      s_last[j] = c > 0 ? g.ep[(size_t)j * k + c - 1] : 0.f;
    }
  }
  __syncthreads();

  const int word_end = (p.row_end + 31) >> 5;
  int cursor = p.row_begin >> 5;  // uniform over the CTA
  if (tid == 0) { s_clk[0] = t_start; s_clk[1] = 0; s_clk[2] = 0; s_clk[3] = 0; }   // diag 2: last stamp, phases a / c / d
  while (true) {
    // ---- lower bounds of this round from the live boards; is any board still filling? ----
    if (tid == 0) { s_bc[2] = 0; s_bc[3] = 0; }
    __syncthreads();
    for (int j = tid; j < C; j += kLbpThreads) {
      float lbv = -INFINITY;
      const int room = k - s_cnt[j];
      if (room <= 0) {
        if (v.srt[j]) {
          lbv = kSet ? s_last[j] : v.ep[(size_t)j * k + k - 1];
        } else if (kSet) {
          lbv = s_min[j];
        } else {
          lbv = INFINITY;
          for (int e = 0; e < k; ++e) lbv = fminf(lbv, v.ep[(size_t)j * k + e]);
        }
      } else {
        atomicAdd(&s_bc[2], 1);
        if (room < kLbpRound) atomicAdd(&s_bc[3], 1);
      }
      s_lb[j] = lbv;
    }
    __syncthreads();
    // While a board is about to fill up every row may matter to it: short rounds refresh the bounds sooner.  A board
    // with room for a whole round appends every own arrival of the round whatever else happens (its count grows by at
    // most one per row), so those rows concern nobody but the board's owner: s_ka.
    const int cap = s_bc[3] > 0 ? kLbpFillRound : kLbpRound;
    for (int j = tid; j < C; j += kLbpThreads) s_ka[j] = (k - s_cnt[j]) >= cap;
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
    if (p.diag == 2 && tid == 0) { const long long t = clock64(); s_clk[1] += t - s_clk[0]; s_clk[0] = t; }
    // ---- (c) relevance: bit s of rel[w] ⇔ row slots[s] beats the bound of some board owned by warp w ----
    // A warp takes 32 slots at a time; lane l tests the boards l, l + 32, … (the boards of WARP l) of every slot and
    // collects warp l's 32 relevance bits in a register: word (l, group) has this one writer, so the slot loop holds
    // no atomics and the loads of several slots are in flight together (one atomicOr per load, each waiting for its
    // L2 round trip, made this phase a quarter of a scan).
    float lbv[4];
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) lbv[gq] = (gq < nb && lane + 32 * gq < C) ? s_lb[lane + 32 * gq] : INFINITY;
    for (int s0 = warp * 32; s0 < n; s0 += kLbpWarps * 32) {
      const int s = s0 + lane;
      int row_s = 0, own_s = 0;
      bool ka = false;
      if (s < n) {
        row_s = slots[s];
        own_s = p.pred[row_s];
        ka = s_ka[own_s] != 0;   // sure to be appended to its own board: nobody else is offered this row
      }
      const uint32_t todo = __ballot_sync(0xffffffffu, s < n && !ka);
      uint32_t bits = 0;
      // two slots per step, their (up to) eight loads issued together and unconditionally: a short-circuit `||` over
      // the column groups made each load wait for the comparison of the one before — four L2 round trips in a row
#pragma unroll 1
      for (int e0 = 0; e0 < 32; e0 += 2) {
        if (!((todo >> e0) & 3u)) continue;
        float val[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool on = (todo >> (e0 + u)) & 1u;
          const int r = __shfl_sync(0xffffffffu, row_s, e0 + u);
          const float* prow = p.rows + (size_t)((on ? r : p.rows_row0) - p.rows_row0) * C;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            val[u][gq] = (on && gq < nb && lane + 32 * gq < C) ? prow[lane + 32 * gq] : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool hit = (val[u][0] > lbv[0]) | (val[u][1] > lbv[1]) | (val[u][2] > lbv[2]) | (val[u][3] > lbv[3]);
          if (hit) bits |= 1u << (e0 + u);
        }
      }
      rel[lane * kRelLd + (s0 >> 5)] = bits;
      __syncwarp();
      if (ka) atomicOr(&rel[(own_s & 31) * kRelLd + (s >> 5)], 1u << (s & 31));
    }
    __syncthreads();
    if (p.diag == 2 && tid == 0) { const long long t = clock64(); s_clk[2] += t - s_clk[0]; s_clk[0] = t; }
    // ---- (d) replay: every warp goes through the words of its relevant rows in index order ----
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
      // Rows another warp owns matter here only if they were REJECTED by their own board (then they are offered
      // to mine), and that is rare — walking them one by one (eight shuffles and a poll per row and warp) made the
      // early part of a scan, where every row is relevant to every warp, issue-bound at ≈1000 clocks per row.  So the
      // word is cut at the rows this warp owns: the verdicts of the foreign rows in between are read by all lanes
      // at once (lane e polls the decision byte of slot e), and only the rejected ones are then walked, in order.
      const bool bit_l = (m >> lane) & 1u;
      const uint32_t mine_mask = __ballot_sync(0xffffffffu, bit_l && (own_e & 31) == warp);
      // at or below the bound of its own board: rejected whatever happened since (its owner does not publish)
      const uint32_t krej_mask = __ballot_sync(0xffffffffu, bit_l && !(pown_e > s_lb[own_e]));

      auto fetch = [&](int e, int& own, float& p_own, int& idx, float& v0, float& v1, float& v2, float& v3) {
        own = __shfl_sync(0xffffffffu, own_e, e);
        p_own = __shfl_sync(0xffffffffu, pown_e, e);
        idx = __shfl_sync(0xffffffffu, row_e, e) + p.idx0;
        v0 = __shfl_sync(0xffffffffu, val_e[0], e); v1 = __shfl_sync(0xffffffffu, val_e[1], e);
        v2 = __shfl_sync(0xffffffffu, val_e[2], e); v3 = __shfl_sync(0xffffffffu, val_e[3], e);
      };
      // :83-101 — a row rejected by its own board is offered to every other board (here: to this warp's boards)
      auto offer = [&](int own, int idx, float v0, float v1, float v2, float v3) -> bool {
        const float p_mine = lane == 0 ? v0 : lane == 1 ? v1 : lane == 2 ? v2 : v3;
        bool need = false;
        if (have_board && my_board != own) {
          const int cj = s_cnt[my_board];
          if (cj < k) {
            v.ep[(size_t)my_board * k + cj] = p_mine;
            v.ei[(size_t)my_board * k + cj] = idx;
            s_cnt[my_board] = cj + 1;
            s_last[my_board] = p_mine;
            if constexpr (kSet) {
              s_min[my_board] = fminf(s_min[my_board], p_mine);
              if (G > 0) s_gm[(size_t)my_board * G + (cj >> 5)] = fminf(s_gm[(size_t)my_board * G + (cj >> 5)], p_mine);
            }
          } else if (s_last[my_board] < p_mine) {
            need = true;
          }
        }
        uint32_t nm = __ballot_sync(0xffffffffu, need);
        const bool admitted = nm != 0;
        while (nm) {
          const int l = __ffs(nm) - 1;
          nm &= nm - 1;
          const float pj = l == 0 ? v0 : l == 1 ? v1 : l == 2 ? v2 : v3;
          if (p.diag == 1 && lane == 0) atomicAdd(&s_dbg[3], 1);
          if constexpr (kSet) {
            const int jb = warp + 32 * l;
            if (G > 0) lb_admit_set_gm(v, jb, pj, idx, p.rank, s_last, s_gm + (size_t)jb * G, G, lane);
            else lb_admit_set(v, jb, pj, idx, p.rank, s_last, lane);
          } else {
            lb_admit(v, warp + 32 * l, pj, idx, p.rank, s_last, lane);
          }
        }
        __syncwarp();
        return admitted;
      };

      uint32_t mm = m;
      while (mm) {
        const uint32_t mine_left = mm & mine_mask;
        const int o = mine_left ? __ffs(mine_left) - 1 : 32;
        const uint32_t seg = mm & ~mine_mask & (o >= 32 ? 0xffffffffu : ((1u << o) - 1u));  // foreign rows before my next one
        // My next own row is fetched and provisionally decided BEFORE the foreign verdicts are waited for: the warps
        // form a chain through the decision bytes (row s is decided once the rows before it are), and what sits
        // between "the last foreign verdict is in" and "my verdict is out" is that chain's step — a poll, two votes
        // and a store instead of the ≈100 dependent instructions of fetch + decide.
        int own = 0, idx = 0, c_own = 0;
        float p_own = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        bool fill = false, acc = false;
        if (o < 32) {
          fetch(o, own, p_own, idx, v0, v1, v2, v3);
          c_own = (int)lbp_lds_u32(last_s + 4u * (uint32_t)(C + own));
          fill = c_own < k;
          acc = fill || __uint_as_float(lbp_lds_u32(last_s + 4u * (uint32_t)own)) < p_own;
        }
        if (seg) {
          uint8_t d = 2;
          const bool poll = ((seg & ~krej_mask) >> lane) & 1u;
          if (p.diag == 1 && poll) atomicAdd(&s_dbg[1], 1);
          while (true) {   // warp-uniform loop (a divergent per-lane spin was 5x slower): lane e looks at slot e's byte
            if (poll) d = (uint8_t)lbp_lds_u8(dec_s + wi * 32 + lane);
            if (__all_sync(0xffffffffu, d != 0)) break;
            if (p.spin_sleep) __nanosleep(p.spin_sleep);
          }
          // Rejected rows are offered to my boards — and nearly always declined: a board that is full but has not
          // admitted anything yet keeps the p of its k-th ARRIVAL as threshold while the only bound that holds for the
          // whole round is its smallest entry, so it finds every rejected row of the round relevant.  Lane e holds
          // row e's probabilities for my boards: all 32 rows are screened at once against the LIVE thresholds (exact
          // at this moment; thresholds change only through an admission, which can lower one — the first admission
          // of a board sorts it — so the rows not yet walked are screened again after every admission), and only
          // the rows that pass are walked, in order.
          uint32_t todo = __ballot_sync(0xffffffffu, ((seg >> lane) & 1u) && d == 2);
          bool offered = false;
          while (todo) {
            bool take = false;
            if ((todo >> lane) & 1u) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                const int b = warp + 32 * gq;
                if (gq < nb && b < C && b != own_e)
                  take = take || s_cnt[b] < k || s_last[b] < val_e[gq];
              }
            }
            uint32_t rej = __ballot_sync(0xffffffffu, take);
            if (rej == 0) break;
            bool admitted = false;
            while (rej && !admitted) {
              const int e = __ffs(rej) - 1;
              rej &= rej - 1;
              todo &= ~((2u << e) - 1u);   // rows up to e are settled
              int own_f, idx_f;
              float p_f, f0, f1, f2, f3;
              fetch(e, own_f, p_f, idx_f, f0, f1, f2, f3);
              if (p.diag == 1 && lane == 0) atomicAdd(&s_dbg[0], 1);
              admitted = offer(own_f, idx_f, f0, f1, f2, f3);
              offered = true;
            }
            if (!admitted) break;   // every row that passed has been walked and nothing changed in between
          }
          mm &= ~seg;
          if (offered && o < 32) {   // an offer may have gone into the board my own row is about to meet
            c_own = (int)lbp_lds_u32(last_s + 4u * (uint32_t)(C + own));
            fill = c_own < k;
            acc = fill || __uint_as_float(lbp_lds_u32(last_s + 4u * (uint32_t)own)) < p_own;
          }
        }
        if (o < 32) {
          // this warp owns the arg-max board: utils/clip_pseudolabels.py:73-82
          mm &= ~(1u << o);
          const int s = wi * 32 + o;
          if (p.diag == 1 && lane == 0) atomicAdd(&s_dbg[0], 1);
          // the verdict first, the bookkeeping after: the warps that wait for it are the critical path
          if (!((krej_mask >> o) & 1u) && lane == 0) lbp_sts_u8(dec_s + s, acc ? 1u : 2u);
          if (p.diag == 5 && lane == 0 && idx < (1 << 21)) p.ts[idx] = clock64();
          if (fill) {
            __syncwarp();
            if (lane == 0) {
              v.ep[(size_t)own * k + c_own] = p_own;
              v.ei[(size_t)own * k + c_own] = idx;
              s_cnt[own] = c_own + 1;
              s_last[own] = p_own;
              if constexpr (kSet) {
                s_min[own] = fminf(s_min[own], p_own);
                if (G > 0) s_gm[(size_t)own * G + (c_own >> 5)] = fminf(s_gm[(size_t)own * G + (c_own >> 5)], p_own);
              }
            }
            __syncwarp();
          } else if (acc) {
            if constexpr (kSet) {
              if (G > 0) lb_admit_set_gm(v, own, p_own, idx, p.rank, s_last, s_gm + (size_t)own * G, G, lane);
              else lb_admit_set(v, own, p_own, idx, p.rank, s_last, lane);
            } else {
              lb_admit(v, own, p_own, idx, p.rank, s_last, lane);
            }
          } else {
            offer(own, idx, v0, v1, v2, v3);
          }
        }
      }
    }
    __syncthreads();
    if (p.diag == 2 && tid == 0) { const long long t = clock64(); s_clk[3] += t - s_clk[0]; s_clk[0] = t; }
  }
  __syncthreads();
  if (p.diag == 2 && tid == 0) {   // phase clocks / 1024 instead of the event counters: collect+bounds, relevance, walk, rounds
    s_dbg[0] = (int32_t)(s_clk[1] >> 10); s_dbg[1] = (int32_t)(s_clk[2] >> 10); s_dbg[3] = (int32_t)(s_clk[3] >> 10); s_dbg[2] = s_dbg[5];
  }
  if (tid == 0) {  // diagnostics, accumulated over the launches of a scan: header words 3..7 =
                   // events walked, decisions waited for, flagged rows, spill admissions, kernel clocks / 1024
    int32_t* hdr = reinterpret_cast<int32_t*>(p.state);
    hdr[3] += s_dbg[0]; hdr[4] += s_dbg[1]; hdr[5] += s_dbg[2]; hdr[6] += s_dbg[3];
    hdr[7] += (int32_t)((clock64() - t_start) >> 10);
  }
  if constexpr (!kSet) {
    for (int i = tid; i < C * k; i += kLbpThreads) { g.ep[i] = v.ep[i]; g.ei[i] = v.ei[i]; }
  }
  for (int j = tid; j < C; j += kLbpThreads) {
    if constexpr (!kSet) g.srt[j] = v.srt[j];
    const int c = s_cnt[j];
    g.cnt[j] = c;
    float lbv = -INFINITY;
    if (c >= k) {
      if (v.srt[j]) {
        lbv = kSet ? s_last[j] : v.ep[(size_t)j * k + k - 1];
      } else if (kSet) {
        lbv = s_min[j];
      } else {
        lbv = INFINITY;
        for (int e = 0; e < k; ++e) lbv = fminf(lbv, v.ep[(size_t)j * k + e]);
      }
    }
    g.lb[j] = lbv;
  }
}

// kSet boards → the list order the reference shows (descending (p, path), utils/clip_pseudolabels.py:78-82): every
// entry of a board that has had an admission counts the entries in front of it (tiles of the board through shared
// memory; keys are distinct) and lands at that position of the scratch copy, which lb_set_copy_kernel moves back.
// Grid (ceil(k / 256), C).
__global__ void __launch_bounds__(256) lb_set_sort_kernel(void* state, int C, int k, const int32_t* __restrict__ rank,
                                                          float* __restrict__ out_p, int32_t* __restrict__ out_i) {
  const LbView v = lb_view(state, C, k);
  const int j = blockIdx.y;
  if (!v.srt[j] || v.cnt[j] < k) return;   // arrival order stands
  __shared__ float t_p[256];
  __shared__ int32_t t_r[256];
  const float* ep = v.ep + (size_t)j * k;
  const int32_t* ei = v.ei + (size_t)j * k;
  const int a = blockIdx.x * 256 + threadIdx.x;
  float pa = 0.f;
  int ia = 0, ra = 0;
  if (a < k) { pa = ep[a]; ia = ei[a]; ra = rank ? rank[ia] : ia; }
  int before = 0;
  for (int t0 = 0; t0 < k; t0 += 256) {
    const int b = t0 + threadIdx.x;
    if (b < k) {
      const int ib = ei[b];
      t_p[threadIdx.x] = ep[b];
      t_r[threadIdx.x] = rank ? rank[ib] : ib;
    }
    __syncthreads();
    const int nt = min(256, k - t0);
    for (int b2 = 0; b2 < nt; ++b2) before += lb_before(t_p[b2], t_r[b2], pa, ra) ? 1 : 0;
    __syncthreads();
  }
  if (a < k) { out_p[(size_t)j * k + before] = pa; out_i[(size_t)j * k + before] = ia; }
}
__global__ void __launch_bounds__(256) lb_set_copy_kernel(void* state, int C, int k, const float* __restrict__ in_p,
                                                          const int32_t* __restrict__ in_i) {
  const LbView v = lb_view(state, C, k);
  const int j = blockIdx.y;
  if (!v.srt[j] || v.cnt[j] < k) return;
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a < k) { v.ep[(size_t)j * k + a] = in_p[(size_t)j * k + a]; v.ei[(size_t)j * k + a] = in_i[(size_t)j * k + a]; }
}
