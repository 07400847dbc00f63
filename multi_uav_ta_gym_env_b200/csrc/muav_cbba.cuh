// CBBA market baseline on the device (planner 7): CBBAReplan.allocate_tasks (TaskAllocation/MarketBased/CBBA_Replan.py:
// 15-69) around CBBA.allocate_tasks (TaskAllocation/MarketBased/CBBA.py:68-324) as the reference's drivers call them:
// max_tasks_per_agent = 1, visibility map, a fresh CBBA(seed + n_replans) per replan (experiments/wps_eval.py:105,134-146,
// escort_eval.py:108-112,149-161).
//
// What makes the reference's result a function of its inputs at all is reproduced here exactly:
//   * the generator: random.Random(seed + n_replans) -- MT19937 seeded by init_by_array([seed]) (CPython
//     Modules/_randommodule.c), shuffle() = reverse Fisher-Yates on _randbelow (top bits of one word, rejection);
//   * the order every auction round starts from, `list(remaining)` of a SET of slot-key strings (CBBA.py:116,128): CPython's
//     str hash (SipHash-1-3, zero key = PYTHONHASHSEED=0) and its set table (8 slots, linear probes 9, perturb shift 5,
//     growth at fill * 5 >= mask * 3 to the first power of two above 4 * used, dummies left by discard) -- oracle/pyset.py
//     is the same restatement in Python, checked against the interpreter (tests/test_pyset.py).
// Fixtures: tests/golden/wps_{hard,commit,escort}_cbba.json.gz, generated from the unmodified reference under
// PYTHONHASHSEED=0; replayed bit-exactly by the oracle, the CPU build of this file and the CUDA path.
//
// The auction is sequential by construction (every bid depends on the paths the previous bids built), so it runs on one
// lane; it is a baseline allocator, not a hot path.
#pragma once
#include "muav_core.cuh"

namespace muav {

constexpr int CBBA_MAX_SLOTS = 128;   // auction slots per call (ERR_NO_SPACE beyond)
constexpr int CBBA_TABLE = 512;       // set table entries (enough for 306 keys)
constexpr int CBBA_PATH = 32;         // slots on one agent's path at a time (ERR_NO_SPACE beyond)
constexpr int CBBA_MAX_BUNDLE = 4;    // max_tasks_per_agent (muav_alloc_opts.max_tasks_per_agent, 1 when 0)

struct CbbaScratch {
  uint32_t* mt;        // [624] MT19937 state
  uint64_t* hash;      // [CBBA_MAX_SLOTS] str hash of the slot key
  double* bid;         // [CBBA_MAX_SLOTS] winning bid
  int16_t* s_tid;      // [CBBA_MAX_SLOTS] task id of the slot
  int16_t* s_kk;       // [CBBA_MAX_SLOTS] ('c' or 'r') << 8 | k
  int16_t* win;        // [CBBA_MAX_SLOTS] winning agent id or -1
  int16_t* table;      // [CBBA_TABLE] set of remaining slots: slot index, -1 empty, -2 dummy
  int16_t* ordered;    // [CBBA_MAX_SLOTS]
  int16_t* live;       // [A] live agent ids, then [A] shuffled copy
  int16_t* bundle;     // [A][CBBA_MAX_BUNDLE] slots held by the agent, in the order they were won
  int16_t* blen;       // [A]
  int16_t* plen;       // [A]
  int16_t* path;       // [A][CBBA_PATH] slot indices
};

MUAV_HD constexpr inline int32_t cbba_scratch_bytes(int A) {
  int b = 624 * 4 + CBBA_MAX_SLOTS * (8 + 8 + 2 + 2 + 2 + 2) + CBBA_TABLE * 2 + A * 2 * (2 + CBBA_MAX_BUNDLE + 1 + 1 + CBBA_PATH) + 64;
  return (b + 15) / 16 * 16;
}

MUAV_HD inline CbbaScratch carve_cbba(char* p, int A) {
  CbbaScratch W;
  W.hash = (uint64_t*)p; p += 8 * CBBA_MAX_SLOTS;
  W.bid = (double*)p; p += 8 * CBBA_MAX_SLOTS;
  W.mt = (uint32_t*)p; p += 4 * 624;
  W.s_tid = (int16_t*)p; p += 2 * CBBA_MAX_SLOTS;
  W.s_kk = (int16_t*)p; p += 2 * CBBA_MAX_SLOTS;
  W.win = (int16_t*)p; p += 2 * CBBA_MAX_SLOTS;
  W.ordered = (int16_t*)p; p += 2 * CBBA_MAX_SLOTS;
  W.table = (int16_t*)p; p += 2 * CBBA_TABLE;
  W.live = (int16_t*)p; p += 2 * 2 * A;
  W.bundle = (int16_t*)p; p += 2 * A * CBBA_MAX_BUNDLE;
  W.blen = (int16_t*)p; p += 2 * A;
  W.plen = (int16_t*)p; p += 2 * A;
  W.path = (int16_t*)p;
  return W;
}

// ---- CPython str hash under PYTHONHASHSEED=0: SipHash-1-3 with a zero key (Python/pyhash.c)
MUAV_HD inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
MUAV_HD inline uint64_t siphash13_zero_key(const unsigned char* s, int n) {
  uint64_t v0 = 0x736F6D6570736575ull, v1 = 0x646F72616E646F6Dull, v2 = 0x6C7967656E657261ull, v3 = 0x7465646279746573ull;
#define MUAV_SIPROUND                                                     \
  v0 += v1; v2 += v3; v1 = rotl64(v1, 13) ^ v0; v3 = rotl64(v3, 16) ^ v2; \
  v0 = rotl64(v0, 32); v2 += v1; v0 += v3; v1 = rotl64(v1, 17) ^ v2;      \
  v3 = rotl64(v3, 21) ^ v0; v2 = rotl64(v2, 32);
  uint64_t b = (uint64_t)n << 56;
  int i = 0;
  for (; n - i >= 8; i += 8) {
    uint64_t m = 0;
    for (int j = 0; j < 8; ++j) m |= (uint64_t)s[i + j] << (8 * j);
    v3 ^= m;
    MUAV_SIPROUND
    v0 ^= m;
  }
  for (int j = 0; i + j < n; ++j) b |= (uint64_t)s[i + j] << (8 * j);
  v3 ^= b;
  MUAV_SIPROUND
  v0 ^= b;
  v2 ^= 0xff;
  MUAV_SIPROUND
  MUAV_SIPROUND
  MUAV_SIPROUND
#undef MUAV_SIPROUND
  uint64_t h = (v0 ^ v1) ^ (v2 ^ v3);
  if (h == ~0ull) h = ~0ull - 1;   // (Py_hash_t)-1 is reserved
  return h;
}
// hash of the slot key f"{tid}#{c}{k}" (CBBA.py:56-64)
MUAV_HD inline uint64_t slot_key_hash(int tid, int kind, int k) {
  unsigned char buf[24];
  int n = 0;
  char tmp[12];
  int m = 0;
  do { tmp[m++] = (char)('0' + tid % 10); tid /= 10; } while (tid > 0);
  while (m > 0) buf[n++] = (unsigned char)tmp[--m];
  buf[n++] = '#';
  buf[n++] = (unsigned char)kind;
  do { tmp[m++] = (char)('0' + k % 10); k /= 10; } while (k > 0);
  while (m > 0) buf[n++] = (unsigned char)tmp[--m];
  return siphash13_zero_key(buf, n);
}

// ---- CPython set (Objects/setobject.c, 3.12) of slot indices
struct CbbaSet {
  int16_t* table;
  const uint64_t* hash;
  int mask, fill, used;
  bool overflow;
  MUAV_HD void clear_table(int size) {
    for (int i = 0; i < size; ++i) table[i] = -1;
    mask = size - 1;
  }
  MUAV_HD void insert_clean(int16_t* t, int msk, int s) const {
    const uint64_t h = hash[s];
    uint64_t perturb = h;
    size_t i = (size_t)h & (size_t)msk;
    for (;;) {
      if (t[i] == -1) { t[i] = (int16_t)s; return; }
      if (i + 9 <= (size_t)msk) {
        for (int j = 1; j <= 9; ++j)
          if (t[i + j] == -1) { t[i + j] = (int16_t)s; return; }
      }
      perturb >>= 5;
      i = (i * 5 + 1 + (size_t)perturb) & (size_t)msk;
    }
  }
  MUAV_HD void resize(int minused) {
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    if (newsize > CBBA_TABLE) { overflow = true; return; }
    // rebuild in place through a stack copy of the live entries in table order
    int16_t old[CBBA_MAX_SLOTS];
    int n = 0;
    for (int i = 0; i <= mask; ++i)
      if (table[i] >= 0 && n < CBBA_MAX_SLOTS) old[n++] = table[i];
    clear_table(newsize);
    for (int i = 0; i < n; ++i) insert_clean(table, mask, old[i]);
    fill = used;
  }
  MUAV_HD void add(int s) {   // keys are distinct: no equality probe needed beyond the slot index
    const uint64_t h = hash[s];
    uint64_t perturb = h;
    size_t i = (size_t)h & (size_t)mask;
    int freeslot = -1;
    for (;;) {
      const int probes = (i + 9 <= (size_t)mask) ? 9 : 0;
      int unused = -1;
      for (int j = 0; j <= probes; ++j) {
        const int16_t e = table[i + j];
        if (e == -1) { unused = (int)(i + j); break; }
        if (e == -2) { if (freeslot < 0) freeslot = (int)(i + j); }
        else if (e == s) return;
      }
      if (unused >= 0) {
        if (freeslot >= 0) { table[freeslot] = (int16_t)s; ++used; return; }
        table[unused] = (int16_t)s;
        ++fill;
        ++used;
        if (fill * 5 >= mask * 3) resize(used > 50000 ? used * 2 : used * 4);
        return;
      }
      perturb >>= 5;
      i = (i * 5 + 1 + (size_t)perturb) & (size_t)mask;
    }
  }
  MUAV_HD void discard(int s) {
    const uint64_t h = hash[s];
    uint64_t perturb = h;
    size_t i = (size_t)h & (size_t)mask;
    for (;;) {
      const int probes = (i + 9 <= (size_t)mask) ? 9 : 0;
      for (int j = 0; j <= probes; ++j) {
        const int16_t e = table[i + j];
        if (e == -1) return;
        if (e == s) { table[i + j] = -2; --used; return; }
      }
      perturb >>= 5;
      i = (i * 5 + 1 + (size_t)perturb) & (size_t)mask;
    }
  }
};

// ---- random.Random(seed): MT19937 (Modules/_randommodule.c)
struct CbbaRng {
  uint32_t* mt;
  int mti;
  MUAV_HD void seed(uint32_t key) {   // init_by_array([key]) -- random.seed(int) for 0 <= int < 2**32
    mt[0] = 19650218u;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1;
    for (int k = 624; k > 0; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key + 0u;   // key[j] + j with one key word: j == 0
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    for (int k = 623; k > 0; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    mti = 624;
  }
  MUAV_HD uint32_t next() {
    if (mti >= 624) {
      for (int kk = 0; kk < 624; ++kk) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
        mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  MUAV_HD int below(int n) {   // _randbelow_with_getrandbits
    int k = 0;
    while ((n >> k) != 0) ++k;
    uint32_t r = next() >> (32 - k);
    while ((int)r >= n) r = next() >> (32 - k);
    return (int)r;
  }
  MUAV_HD void shuffle(int16_t* x, int n) {
    for (int i = n - 1; i >= 1; --i) {
      const int j = below(i + 1);
      const int16_t t = x[i];
      x[i] = x[j];
      x[j] = t;
    }
  }
};

struct CbbaCtx {
  Sim* S;
  CbbaScratch W;
  double max_dist, makespan;
  // calculate_task_score (CBBA.py:288-309)
  MUAV_HD double task_score(int a, int tid, double px, double py, double time) const {
    const View& V = S->V;
    const int k = tid - 1;
    const double dist = norm2(px - V.k_posx()[k], py - V.k_posy()[k]);
    const int ti = V.k_type()[k];
    double quality = S->cap(a, ti);
    if (is_coalition(*S, k)) quality = dmax(quality, 1.0);
    const double sp = S->speed_of(a);
    const double speed = dmax(sp != 0.0 ? sp : 1.0, 1e-6);
    time = time + ddiv(dist, speed);
    const int dl = V.k_deadline()[k];
    if (dl >= 0 && time > (double)dl) return -50.0;
    const double base = ddiv(-2.5 * dist, dmax(max_dist, 1.0)) + 160.0 * quality;
    if (time < makespan) return base + 2.0 * (makespan - time);
    return base - 2.0 * (time - makespan);
  }
  // _score_mixed_path (CBBA.py:262-272) of agent a's path with `tid` inserted at position `at` (at < 0: the path alone)
  MUAV_HD double score_path(int a, int row, int tid, int at) const {
    const View& V = S->V;
    double score = 0.0;
    double px = V.a_posx()[a], py = V.a_posy()[a];
    double time = V.a_nft()[a];
    const double sp = S->speed_of(a);
    const double speed = dmax(sp != 0.0 ? sp : 1.0, 1e-6);
    const int n = W.plen[row];
    const int total = n + (at >= 0 ? 1 : 0);
    for (int i = 0; i < total; ++i) {
      int t;
      if (at >= 0 && i == at) t = tid;
      else t = W.s_tid[W.path[row * CBBA_PATH + (at >= 0 && i > at ? i - 1 : i)]];
      score += task_score(a, t, px, py, time);
      const int k = t - 1;
      const double dist = norm2(px - V.k_posx()[k], py - V.k_posy()[k]);
      px = V.k_posx()[k];
      py = V.k_posy()[k];
      time += ddiv(dist, speed) + (double)S->C().duration[V.k_type()[k]];
    }
    return score;
  }
  MUAV_HD double total_time(int a, int row) const {
    const View& V = S->V;
    double px = V.a_posx()[a], py = V.a_posy()[a];
    double time = V.a_nft()[a];
    const double sp = S->speed_of(a);
    const double speed = dmax(sp != 0.0 ? sp : 1.0, 1e-6);
    for (int i = 0; i < W.plen[row]; ++i) {
      const int k = W.s_tid[W.path[row * CBBA_PATH + i]] - 1;
      const double dist = norm2(px - V.k_posx()[k], py - V.k_posy()[k]);
      px = V.k_posx()[k];
      py = V.k_posy()[k];
      time += ddiv(dist, speed) + (double)S->C().duration[V.k_type()[k]];
    }
    return time;
  }
  MUAV_HD void path_remove(int row, int s) {
    int n = W.plen[row];
    int16_t* p = W.path + row * CBBA_PATH;
    for (int i = 0; i < n; ++i)
      if (p[i] == s) {
        for (int j = i; j + 1 < n; ++j) p[j] = p[j + 1];
        W.plen[row] = (int16_t)(n - 1);
        return;
      }
  }
};

MUAV_HD MUAV_NI_A inline int cbba_allocate(Sim& S, const muav_alloc_opts& O, int e, int16_t* out_agent, int16_t* out_tid, int lane,
                                           int nlanes) {
  View& V = S.V;
  const int A = V.lay().D.A;
  const int t = HIv(T);
  int32_t* ctrl = (int32_t*)(S.scratch + cbba_scratch_bytes(A) - 16);
  if (lane == 0) {
    int n_pairs = 0;
    HIv(N_CALLS) += 1;
    if (O.d_n_bundle_pairs) O.d_n_bundle_pairs[e] = 0;
    const bool ev_hit = (HIv(EV_TAGMASK) & O.event_mask) != 0;
    const int interval = O.replan_interval > 0 ? O.replan_interval : 1;
    bool go;
    if (O.mode == 1) go = (t - HIv(LAST_PLAN_STEP)) >= interval || ev_hit;
    else if (O.mode == 2) go = t == 0 || (t % interval) == 0 || ev_hit;
    else go = O.mode == 3;
    if (go) {
      HIv(LAST_PLAN_STEP) = t;
      HIv(N_REPLANS) += 1;
      CbbaCtx X;
      X.S = &S;
      X.W = carve_cbba(S.scratch, A);
      X.max_dist = O.max_coord;
      X.makespan = 0.0;
      CbbaScratch& W = X.W;
      const uint8_t* reserved = O.d_reserved ? O.d_reserved + (size_t)e * A : nullptr;
      int n_live = 0;
      for (int a = 0; a < A; ++a)
        if (V.a_state()[a] != -1 && !(reserved && reserved[a])) W.live[n_live++] = (int16_t)a;
      // expand_slot_keys over _open_tasks (CBBA.py:46-65, paper_eval.py:96-101)
      int n_slot = 0, n_tasks_arg = 0;
      const int n_tasks = HIv(N_TASKS);
      const int KWn = (n_tasks + 31) >> 5;
      const int IC = V.lay().D.IC;
      const int32_t* order = O.d_task_order ? O.d_task_order + (size_t)e * IC : nullptr;   // the caller's `tasks` argument
      int wd = 0;
      uint32_t bits = order ? 0u : (KWn > 0 ? V.open_mask()[0] : 0u);
      for (int it = 0;; ++it) {
          int k;
          if (order) {
            if (it >= IC) break;
            k = order[it];
            if (k < 0) break;
            if (k >= n_tasks) continue;
          } else {
            while (!bits && ++wd < KWn) bits = V.open_mask()[wd];
            if (!bits) break;
            k = (wd << 5) + ctz32(bits);
            bits &= bits - 1;
          }
          if (V.k_status()[k] == 2) continue;
          ++n_tasks_arg;
          const double rem = residual_demand(S, k);
          if (!(rem > 0)) continue;
          const bool coal = is_coalition(S, k);
          int ns = coal ? (int)ceil(rem) : (int)ceil(dmin(rem, 4.0));
          if (!coal && ns < 1) ns = 1;
          for (int j = 0; j < ns; ++j) {
            if (n_slot >= CBBA_MAX_SLOTS) { HIv(ERRFLAGS) |= ERR_NO_SPACE; break; }
            W.s_tid[n_slot] = (int16_t)(k + 1);
            W.s_kk[n_slot] = (int16_t)(((coal ? 'c' : 'r') << 8) | j);
            W.hash[n_slot] = slot_key_hash(k + 1, coal ? 'c' : 'r', j);
            W.win[n_slot] = -1;
            W.bid[n_slot] = -INFINITY;
            ++n_slot;
          }
      }
      if (n_live > 0 && n_tasks_arg > 0 && n_slot > 0) {
        CbbaSet R;
        R.table = W.table;
        R.hash = W.hash;
        R.fill = R.used = 0;
        R.overflow = false;
        R.clear_table(8);
        for (int s = 0; s < n_slot; ++s) R.add(s);
        for (int i = 0; i < n_live; ++i) { W.blen[i] = 0; W.plen[i] = 0; }
        int MB = O.max_tasks_per_agent > 1 ? O.max_tasks_per_agent : 1;
        if (MB > CBBA_MAX_BUNDLE) MB = CBBA_MAX_BUNDLE;
        // bundle helpers: position of slot s / of a slot of task tid in the bundle of a row, removal keeps the order
        auto b_find = [&](int row, int s2) { for (int j = 0; j < W.blen[row]; ++j) if (W.bundle[row * CBBA_MAX_BUNDLE + j] == s2) return j; return -1; };
        auto b_owns = [&](int row, int tid2) { for (int j = 0; j < W.blen[row]; ++j) if (W.s_tid[W.bundle[row * CBBA_MAX_BUNDLE + j]] == tid2) return true; return false; };
        auto b_remove = [&](int row, int s2) { const int j0 = b_find(row, s2); if (j0 < 0) return; for (int j = j0; j + 1 < W.blen[row]; ++j) W.bundle[row * CBBA_MAX_BUNDLE + j] = W.bundle[row * CBBA_MAX_BUNDLE + j + 1]; W.blen[row] -= 1; };
        CbbaRng G;
        G.mt = W.mt;
        const uint32_t seed0 = O.d_cbba_seed ? (uint32_t)O.d_cbba_seed[e] : 0u;
        G.seed(seed0 + (uint32_t)HIv(N_REPLANS));
        const int max_iters = 2 * n_slot > 8 ? 2 * n_slot : 8;
        int16_t* order_a = W.live + A;
        for (int it = 0; it < max_iters && R.used > 0; ++it) {
          bool changed = false;
          int n_ord = 0;
          for (int i = 0; i <= R.mask; ++i)
            if (R.table[i] >= 0) W.ordered[n_ord++] = R.table[i];
          G.shuffle(W.ordered, n_ord);
          for (int i = 0; i < n_live; ++i) order_a[i] = (int16_t)i;   // rows of `live`
          G.shuffle(order_a, n_live);
          for (int oi = 0; oi < n_ord; ++oi) {
            const int s = W.ordered[oi];
            const int tid = W.s_tid[s];
            const int k = tid - 1;
            const int el = V.k_elig()[k];
            const bool coal = is_coalition(S, k);
            for (int ai = 0; ai < n_live; ++ai) {
              const int row = order_a[ai];
              const int a = W.live[row];
              // agent_eligible (CBBA.py:27-43)
              if (O.use_visibility && !S.known_bit(a, k)) continue;
              if (el != 0 && !((el >> V.a_type()[a]) & 1)) continue;
              if (S.qfind(a, tid) >= 0) continue;
              if (!coal && !(S.cap(a, V.k_type()[k]) > 0)) continue;
              if (b_owns(row, tid)) continue;              // task.id in owned_tasks[agent] (also: this slot is in the bundle)
              if (W.blen[row] >= MB) continue;             // bundle full
              // calculate_bid (CBBA.py:216-225)
              double best = -INFINITY;
              const int n = W.plen[row];
              for (int at = 0; at <= n; ++at) {
                const double sc = X.score_path(a, row, tid, at);
                if (sc > best) best = sc;
              }
              const double bid = best - X.score_path(a, row, 0, -1);
              if (bid <= W.bid[s]) continue;
              changed = true;
              const int prev = W.win[s];
              if (prev >= 0) {
                int prow = 0;
                for (int i = 0; i < n_live; ++i)
                  if (W.live[i] == prev) prow = i;
                X.path_remove(prow, s);
                b_remove(prow, s);   // owned_tasks is the set of the bundle's tasks
              }
              W.win[s] = (int16_t)a;
              W.bid[s] = bid;
              // determine_insertion_point (:229-237)
              double mx = -INFINITY;
              int ins = 0;
              const int n2 = W.plen[row];
              for (int at = 0; at <= n2; ++at) {
                const double sc = X.score_path(a, row, tid, at);
                if (sc > mx) { mx = sc; ins = at; }
              }
              if (n2 >= CBBA_PATH) { HIv(ERRFLAGS) |= ERR_NO_SPACE; continue; }
              int16_t* p = W.path + row * CBBA_PATH;
              for (int j = n2; j > ins; --j) p[j] = p[j - 1];
              p[ins] = (int16_t)s;
              W.plen[row] = (int16_t)(n2 + 1);
            }
          }
          if (!changed) break;
          // consensus (:159-190): slots in key-list order
          for (int s = 0; s < n_slot; ++s) {
            const int winner = W.win[s];
            if (winner < 0) continue;
            const int tid = W.s_tid[s];
            int wrow = 0;
            for (int i = 0; i < n_live; ++i) {
              if (W.live[i] == winner) { wrow = i; continue; }
              if (b_find(i, s) >= 0) {
                b_remove(i, s);
                X.path_remove(i, s);
              }
            }
            if (b_find(wrow, s) < 0) {
              if (b_owns(wrow, tid) || W.blen[wrow] >= MB) {   // another slot of the task already owned, or the bundle is full
                W.win[s] = -1;
                W.bid[s] = -INFINITY;
                X.path_remove(wrow, s);
                continue;
              }
              W.bundle[wrow * CBBA_MAX_BUNDLE + W.blen[wrow]] = (int16_t)s;
              W.blen[wrow] += 1;
              R.discard(s);
            }
          }
          double mk = 0.0;
          for (int i = 0; i < n_live; ++i) {
            const double tt = X.total_time(W.live[i], i);
            if (i == 0 || tt > mk) mk = tt;
          }
          X.makespan = mk;
        }
        if (R.overflow) HIv(ERRFLAGS) |= ERR_NO_SPACE;
        // the step's pairs: the first task of every bundle (_apply_assign keeps an agent's first pair); with bundles the
        // whole plan goes to d_bundle_pairs in the order allocate_tasks returns it (CBBA.py:192-204)
        int nb = 0;
        int32_t* bp = (MB > 1 && O.d_bundle_pairs) ? O.d_bundle_pairs + (size_t)e * A * O.max_tasks_per_agent : nullptr;
        for (int i = 0; i < n_live; ++i) {
          if (W.blen[i] > 0) {
            out_agent[n_pairs] = W.live[i];
            out_tid[n_pairs] = W.s_tid[W.bundle[i * CBBA_MAX_BUNDLE]];
            ++n_pairs;
          }
          for (int j = 0; j < W.blen[i]; ++j) {
            if (bp) bp[nb] = ((int)W.live[i] << 16) | (int)W.s_tid[W.bundle[i * CBBA_MAX_BUNDLE + j]];
            ++nb;
          }
        }
        if (MB > 1 && O.d_n_bundle_pairs) O.d_n_bundle_pairs[e] = nb;
      }
    }
    ctrl[0] = n_pairs;
  }
  MUAV_WARP_SYNC();
  const int n_pairs = ctrl[0];
  MUAV_WARP_SYNC();
  return n_pairs;
}

}  // namespace muav
