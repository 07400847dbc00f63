// Att-Pair scorer forward on the 5th-generation tensor cores (tcgen05 + TMEM), same function as att_pair_kernel in
// muav_scorer.cu (AttPairNet, TaskAllocation/Hybrid/PairCostHybrid.py:89-151, scores = tanh(logits) * clamp * edge_valid,
// :266-278), with the context term of AttContextPairNet (ContextPairHybrid.py:81-151) when the parameters have one;
// att_tc_kernel<true> is the AttCommitNet forward (AttentionCommit.py:68-100: the projections, two encoder layers, the
// priority / commit sigmoid heads) on the same machinery.
//
// One CTA per SM owns all 512 TMEM columns.  A pass packs the live tokens of up to eight environments into the 128
// rows of an M = 128 MMA (agents first, then tasks; row = TMEM lane).  Every linear layer is D[128 x N] = A[128 x K] *
// W[N x K]^T with
//   * A (the activations) in TMEM, written there by the previous layer's epilogue with tcgen05.st -- each worker thread
//     owns one token row, so LayerNorm, residuals and ReLU are thread-local and activations never pass through shared
//     memory;
//   * W streamed from L2 into a four-slot shared-memory ring by a producer thread (cp.async.bulk, 1-D TMA), packed by
//     muav_att_pair_tc_pack in the canonical K-major no-swizzle layout the MMA reads;
//   * 3xTF32: x = hi + lo with hi = the upper 19 bits; hi*hi + hi*lo + lo*hi accumulate in fp32 (the dropped lo*lo term
//     is 2^-22 relative), which keeps the 2e-5 score tolerance of the fp32 kernel.
// Token-type specific layers (projections, cross-attention in/out, first pair-head layer) run both weight sets over all
// rows into different TMEM columns; the epilogue of a row reads the columns of its type.
// Attention reads q from TMEM and k / v from a token-major shared-memory tile; the pair head is two more MMAs over
// tiles of 128 valid (agent, task) pairs.
//
// Roles: warps 0-7 workers (row = 32 * (warp & 3) + lane, warps w and w + 4 split the columns), warp 8 lane 0 issues the
// MMAs, warp 9 lane 0 streams the weights.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/muav.h"

namespace muav_tc {

constexpr int D = 64;
constexpr int HD = 16;
constexpr int TF = 13, AF = 12;
constexpr int ROWS = 128;
constexpr int NWORK = 256;
constexpr int NT = 320;
constexpr int NSLOT = 4;
constexpr int SLOT_BYTES = 32768;
constexpr int KV_STRIDE = 132;  // floats per token row of the k | v tile (conflict-free float4 rows)
constexpr int ZG_STRIDE = 68;   // floats per token row of the z and g tiles of the pair head
constexpr int KVG_BYTES = (ROWS * KV_STRIDE > 2 * ROWS * ZG_STRIDE ? ROWS * KV_STRIDE : 2 * ROWS * ZG_STRIDE) * 4;
constexpr int PLIST_MAX = 4096;
constexpr int DYN_SMEM = NSLOT * SLOT_BYTES + KVG_BYTES + PLIST_MAX * 2;
constexpr int GMAX = 8;    // environments per pass
constexpr int GLIST = 16;  // environments per CTA

// TMEM columns
constexpr int A0_HI = 0, A0_LO = 64;        // layer input x (hi / lo), 64 features
constexpr int F_HI = 256, F_LO = 272;       // raw features, K padded to 16
constexpr int AO_HI = 320, AO_LO = 384;     // self-attention output
constexpr int H_HI = 256, H_LO = 384;       // feed-forward hidden, 128 features
constexpr int AO2_HI = 192, AO2_LO = 256;   // cross-attention output
constexpr int U_HI = 256, U_LO = 320;       // pair tile: a*t products, then the first hidden layer
constexpr int P1_D = 384, P2_D = 448;       // pair tile accumulators
constexpr int U2_HI = 0, U2_LO = 64;        // second pair tile of a round (x and the head1 accumulators are dead by then)
constexpr int P1B_D = 128, P2B_D = 192;

// ---- weight chunks in consumption order (one ring slot each).  A chunk holds all N output features of its layer for a
// slice of K (so that one MMA covers the whole N: fewer, larger instructions); token-type specific layers that share
// their A operand are one layer with the two weight matrices stacked along N.
struct ChunkDesc {
  uint32_t off;   // float offset in the packed buffer
  uint16_t nc, kc;
  uint16_t a_hi, a_lo, d_col;
  uint8_t acc, last;
};
constexpr int NCHUNK = 24;
constexpr int CH_PAIR1 = 22, CH_PAIR2 = 23;
#define CH(off, nc, kc, ahi, alo, d, acc, last) \
  ChunkDesc { off, nc, kc, ahi, alo, d, acc, last }
constexpr uint32_t S_PROJ = 128 * 16 * 2, S_IN = 192 * 16 * 2, S_64 = 64 * 64 * 2, S_128 = 128 * 32 * 2, S_H2 = 32 * 64 * 2;
constexpr uint32_t O_IN = S_PROJ, O_OUT = O_IN + 4 * S_IN, O_L1 = O_OUT + S_64, O_L2 = O_L1 + 2 * S_128, O_XA = O_L2 + 2 * S_64,
                   O_XB = O_XA + 4 * S_IN, O_XO = O_XB + 4 * S_IN, O_H1 = O_XO + 2 * S_128, O_P1 = O_H1 + 2 * S_128,
                   O_P2 = O_P1 + S_64;
__constant__ ChunkDesc c_chunks[NCHUNK] = {
    CH(0, 128, 16, F_HI, F_LO, 128, 0, 1),                              // 0 agent_proj | task_proj
    CH(O_IN, 192, 16, A0_HI, A0_LO, 128, 0, 0),                         // 1-4 encoder in_proj (q | k | v), K slices of 16
    CH(O_IN + S_IN, 192, 16, A0_HI + 16, A0_LO + 16, 128, 1, 0),
    CH(O_IN + 2 * S_IN, 192, 16, A0_HI + 32, A0_LO + 32, 128, 1, 0),
    CH(O_IN + 3 * S_IN, 192, 16, A0_HI + 48, A0_LO + 48, 128, 1, 1),
    CH(O_OUT, 64, 64, AO_HI, AO_LO, 448, 0, 1),                         // 5 encoder out_proj
    CH(O_L1, 128, 32, A0_HI, A0_LO, 128, 0, 0),                         // 6-7 linear1, K slices of 32
    CH(O_L1 + S_128, 128, 32, A0_HI + 32, A0_LO + 32, 128, 1, 1),
    CH(O_L2, 64, 64, H_HI, H_LO, 128, 0, 0),                            // 8-9 linear2, K slices of 64
    CH(O_L2 + S_64, 64, 64, H_HI + 64, H_LO + 64, 128, 1, 1),
    CH(O_XA, 192, 16, A0_HI, A0_LO, 128, 0, 0),                         // 10-13 cross_a2t in_proj
    CH(O_XA + S_IN, 192, 16, A0_HI + 16, A0_LO + 16, 128, 1, 0),
    CH(O_XA + 2 * S_IN, 192, 16, A0_HI + 32, A0_LO + 32, 128, 1, 0),
    CH(O_XA + 3 * S_IN, 192, 16, A0_HI + 48, A0_LO + 48, 128, 1, 0),
    CH(O_XB, 192, 16, A0_HI, A0_LO, 320, 0, 0),                         // 14-17 cross_t2a in_proj
    CH(O_XB + S_IN, 192, 16, A0_HI + 16, A0_LO + 16, 320, 1, 0),
    CH(O_XB + 2 * S_IN, 192, 16, A0_HI + 32, A0_LO + 32, 320, 1, 0),
    CH(O_XB + 3 * S_IN, 192, 16, A0_HI + 48, A0_LO + 48, 320, 1, 1),
    CH(O_XO, 128, 32, AO2_HI, AO2_LO, 384, 0, 0),                       // 18-19 cross_a2t | cross_t2a out_proj
    CH(O_XO + S_128, 128, 32, AO2_HI + 32, AO2_LO + 32, 384, 1, 1),
    CH(O_H1, 128, 32, A0_HI, A0_LO, 128, 0, 0),                         // 20-21 head1 agent block | task block
    CH(O_H1 + S_128, 128, 32, A0_HI + 32, A0_LO + 32, 128, 1, 1),
    CH(O_P1, 64, 64, U_HI, U_LO, 384, 0, 1),                            // 22 head1 product block (per pair tile)
    CH(O_P2, 32, 64, U_HI, U_LO, 448, 0, 1),                            // 23 head2 (per pair tile)
};
constexpr uint32_t TCW_FLOATS = O_P2 + S_H2;

// AttCommitNet (AttentionCommit.py:68-100): the projections and two encoder layers, no cross attention / pair head
constexpr int NCHUNK_COMMIT = 19;
constexpr uint32_t S_ENC = 4 * S_IN + S_64 + 2 * S_128 + 2 * S_64;   // one encoder layer
#define ENC_CHUNKS(o)                                                         \
  CH((o), 192, 16, A0_HI, A0_LO, 128, 0, 0),                                  \
  CH((o) + S_IN, 192, 16, A0_HI + 16, A0_LO + 16, 128, 1, 0),                 \
  CH((o) + 2 * S_IN, 192, 16, A0_HI + 32, A0_LO + 32, 128, 1, 0),             \
  CH((o) + 3 * S_IN, 192, 16, A0_HI + 48, A0_LO + 48, 128, 1, 1),             \
  CH((o) + 4 * S_IN, 64, 64, AO_HI, AO_LO, 448, 0, 1),                        \
  CH((o) + 4 * S_IN + S_64, 128, 32, A0_HI, A0_LO, 128, 0, 0),                \
  CH((o) + 4 * S_IN + S_64 + S_128, 128, 32, A0_HI + 32, A0_LO + 32, 128, 1, 1), \
  CH((o) + 4 * S_IN + S_64 + 2 * S_128, 64, 64, H_HI, H_LO, 128, 0, 0),       \
  CH((o) + 4 * S_IN + 2 * S_64 + 2 * S_128, 64, 64, H_HI + 64, H_LO + 64, 128, 1, 1)
__constant__ ChunkDesc c_chunks_commit[NCHUNK_COMMIT] = {
    CH(0, 128, 16, F_HI, F_LO, 128, 0, 1),   // agent_proj | task_proj
    ENC_CHUNKS(S_PROJ),
    ENC_CHUNKS(S_PROJ + S_ENC),
};
constexpr uint32_t TCW_FLOATS_COMMIT = S_PROJ + 2 * S_ENC;
template <bool COMMIT>
__device__ __forceinline__ const ChunkDesc* chunk_table() {
  return COMMIT ? c_chunks_commit : c_chunks;
}

// ---- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// bounded wait (~10 s of SM clocks): a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(phase)
        : "memory");
    if (!done && clock64() - t0 > 20000000000LL) __trap();
  } while (!done);
}
// the same with a back-off between polls (waiters that must not steal issue slots from the MMA thread)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(phase)
        : "memory");
    if (done) break;
    __nanosleep(20);
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// workers: one lane polls the accumulator barrier, the warp follows
__device__ __forceinline__ void wait_d(uint64_t* bar, uint32_t& phase, int lane) {
  if (lane == 0) mbar_wait_relaxed(bar, phase);
  __syncwarp();
  phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 16 consecutive columns of this thread's TMEM lane (issue only; tc_wait_ld() before the values are used)
__device__ __forceinline__ void tmem_ld16(uint32_t a, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(a)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t a, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
// x = hi + lo exactly; hi has the 10 mantissa bits TF32 keeps
__device__ __forceinline__ void split16(const float* x, float* hi, float* lo) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hi[i] = __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
    lo[i] = x[i] - hi[i];
  }
}
// store 16 features of this thread's row as an MMA A operand (hi and lo planes)
__device__ __forceinline__ void st_operand16(uint32_t tl, int col_hi, int col_lo, int c, const float* x) {
  float hi[16], lo[16];
  split16(x, hi, lo);
  tmem_st16(tl + col_hi + c, hi);
  tmem_st16(tl + col_lo + c, lo);
}
// 16 accumulator columns of this thread's row from the column block of its token type.
// wtype (warp-uniform): 0 = every row of the warp is an agent row, 1 = every row a task row, 2 = mixed
__device__ __forceinline__ void ld_variant16(uint32_t tl, int col_agent, int col_task, int c, bool is_agent, int wtype, float* v) {
  if (wtype == 0) {
    tmem_ld16(tl + col_agent + c, v);
  } else if (wtype == 1) {
    tmem_ld16(tl + col_task + c, v);
  } else {
    float b[16];
    tmem_ld16(tl + col_agent + c, v);
    tmem_ld16(tl + col_task + c, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = is_agent ? v[i] : b[i];
  }
}

// shared-memory matrix descriptor: no swizzle, K-major (validated by tools/tc_probe)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: fp32 accumulate, TF32 x TF32, both K-major
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the MMAs of one weight chunk: hi*hi + hi*lo + lo*hi over K in steps of 8
template <int KC>
__device__ __forceinline__ void issue_chunk_k(uint32_t tm, const ChunkDesc& c, uint32_t b0) {
  const uint32_t idesc = make_idesc(ROWS, c.nc);
  const uint32_t lbo = (uint32_t)c.nc * 16u;
  const uint32_t plane16 = ((uint32_t)c.nc * KC * 4u) >> 4;   // lo plane, in 16-byte units
  const uint32_t step16 = (2u * lbo) >> 4;                    // one K = 8 step
  const uint64_t desc0 = make_desc(b0, lbo, 128);
  const uint32_t d = tm + c.d_col;
  uint32_t acc = c.acc;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a = tm + (term == 2 ? c.a_lo : c.a_hi);
    const uint64_t db = desc0 + (term == 1 ? plane16 : 0u);
#pragma unroll
    for (int k8 = 0; k8 < KC / 8; ++k8) {
      mma_ts(d, a + k8 * 8, db + (uint64_t)(k8 * step16), idesc, acc);
      acc = 1;
    }
  }
}
__device__ __forceinline__ void issue_chunk(uint32_t tm, const ChunkDesc& c, const unsigned char* slot) {
  const uint32_t b0 = smem_u32(slot);
  if (c.kc == 16) issue_chunk_k<16>(tm, c, b0);
  else if (c.kc == 32) issue_chunk_k<32>(tm, c, b0);
  else issue_chunk_k<64>(tm, c, b0);
}

constexpr int MAX_ENC = 2;
// parameter offsets of both networks in one form (filled from muav_attpair_offsets / muav_attcommit_offsets)
struct TcOffsets {
  int32_t agent_proj_w, agent_proj_b, task_proj_w, task_proj_b, type_embed;
  int32_t enc_in_w[MAX_ENC], enc_in_b[MAX_ENC], enc_out_w[MAX_ENC], enc_out_b[MAX_ENC], enc_l1_w[MAX_ENC], enc_l1_b[MAX_ENC],
      enc_l2_w[MAX_ENC], enc_l2_b[MAX_ENC], enc_n1_w[MAX_ENC], enc_n1_b[MAX_ENC], enc_n2_w[MAX_ENC], enc_n2_b[MAX_ENC];
  int32_t a2t_in_w, a2t_in_b, a2t_out_w, a2t_out_b, t2a_in_w, t2a_in_b, t2a_out_w, t2a_out_b;
  int32_t head1_w, head1_b, head2_w, head2_b, head3_w, head3_b;
  int32_t ctx_proj_w, ctx_proj_b, has_context;
  int32_t priority_w, priority_b, commit_w, commit_b;   // AttCommitNet heads
};
struct Params {
  const float* w;
  TcOffsets o;
  const float* tcw;
  const float* task_feats;
  const uint8_t* task_mask;
  const float* agent_feats;
  const uint8_t* agent_mask;
  const float* edge_valid;
  const float* context;  // [E, 8] (AttContextPairNet) or NULL
  const int32_t* env_idx;
  const uint8_t* need;
  float* scores;
  float* pri;   // AttCommitNet: priorities [E, max_tasks], commit gates [E, max_agents]
  float* com;
  int n, max_tasks, max_agents;
  float clamp;
  float* dbg;   // development: [stage][128][64] dump of the layer inputs of CTA 0's first pass, or NULL
};

struct Seg {
  int e, abase, tbase, na, nt;
};
#define SEG_NONE 0xFF

struct PickArgs {
  const uint8_t* need;
  const int32_t* env_idx;
  int n;
};
// Launch slots b in [0, n) map to environments e = env_idx ? env_idx[b] : b; a slot counts when need == NULL or
// need[e] != 0 (same rule as muav_scorer.cu).  scan_envs: every thread counts its chunk of slots once; select_envs
// writes the environments of rank [r0, r0 + m) to s_env.
struct EnvScan {
  int lo, hi, cnt, excl, total;
};
__device__ EnvScan scan_envs(const PickArgs P, int* s_scan) {
  const int tid = threadIdx.x;
  EnvScan S;
  if (!P.need) {
    S.lo = S.hi = S.cnt = S.excl = 0;
    S.total = P.n;
    return S;
  }
  const int chunk = (P.n + NT - 1) / NT;
  S.lo = min(P.n, tid * chunk);
  S.hi = min(P.n, S.lo + chunk);
  int cnt = 0;
  for (int b = S.lo; b < S.hi; ++b) cnt += P.need[P.env_idx ? P.env_idx[b] : b] != 0;
  const int lane = tid & 31, wid = tid >> 5;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < NT / 32; ++w) { const int v = s_scan[w]; s_scan[w] = run; run += v; }
    s_scan[NT / 32] = run;
  }
  __syncthreads();
  S.cnt = cnt;
  S.excl = s_scan[wid] + incl - cnt;
  S.total = s_scan[NT / 32];
  return S;
}
__device__ void select_envs(const PickArgs P, const EnvScan S, int r0, int m, int* s_env) {
  const int tid = threadIdx.x;
  if (!P.need) {
    if (tid < m) s_env[tid] = P.env_idx ? P.env_idx[r0 + tid] : r0 + tid;
    return;
  }
  if (S.cnt > 0 && S.excl < r0 + m && S.excl + S.cnt > r0) {
    int r = S.excl;
    for (int b = S.lo; b < S.hi; ++b) {
      const int e = P.env_idx ? P.env_idx[b] : b;
      if (P.need[e] != 0) {
        if (r >= r0 && r < r0 + m) s_env[r - r0] = e;
        ++r;
      }
    }
  }
}

// development dump of this thread's 32 columns of x (A0 planes); compiled in only with -DMUAV_TC_DEBUG (tools/tc_scorer_check.py)
__device__ __forceinline__ void dump_a0(const Params& P, int stage, uint32_t tl, int row, int half) {
#if !defined(MUAV_TC_DEBUG)
  return;
#endif
  if (!P.dbg || blockIdx.x != 0) return;
  for (int g = 0; g < 2; ++g) {
    const int c = half * 32 + g * 16;
    float h[16], l[16];
    tmem_ld16(tl + A0_HI + c, h);
    tmem_ld16(tl + A0_LO + c, l);
    for (int i = 0; i < 16; ++i) P.dbg[((size_t)stage * ROWS + row) * D + c + i] = h[i] + l[i];
  }
}

// softmax(q k^T / 4) v for the two heads of this thread, keys = the task / agent tokens of the row's environment
// (cross: agents attend tasks, tasks attend agents).  q from TMEM, k | v from the shared tile, output to TMEM as the
// A operand of the out-projection.
__device__ __forceinline__ void attention(uint32_t tl, const float* __restrict__ kv, int half, bool on, const Seg sg,
                                          bool cross, bool is_agent, int wtype, int qcol_agent, int qcol_task,
                                          const float* __restrict__ qb_agent, const float* __restrict__ qb_task, int ao_hi,
                                          int ao_lo) {
  int k0[2] = {sg.abase, sg.tbase};
  int k1[2] = {sg.abase + sg.na, sg.tbase + sg.nt};
  if (cross) {
    if (is_agent) k1[0] = k0[0];
    else k1[1] = k0[1];
  }
  const float* qb = is_agent ? qb_agent : qb_task;
  // the thread's two heads (features [32 half, 32 half + 32)) run side by side: two independent softmax chains, and
  // the 16-term dot products as four partial sums each
  float q[2 * HD], acc[2 * HD];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) ld_variant16(tl, qcol_agent, qcol_task, (half * 2 + hh) * HD, is_agent, wtype, q + hh * HD);
#pragma unroll
  for (int d = 0; d < 2 * HD; ++d) {
    q[d] = (q[d] + __ldg(&qb[half * 2 * HD + d])) * (0.25f * 1.44269504088896340736f);   // scores in log2 units
    acc[d] = 0.0f;
  }
  if (on) {
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.0f, 0.0f};
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
      for (int j = k0[rr]; j < k1[rr]; ++j) {
        const float4* kr = (const float4*)&kv[j * KV_STRIDE + half * 2 * HD];
        const float4* vr = (const float4*)&kv[j * KV_STRIDE + D + half * 2 * HD];
        float s[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float part[4];
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 kk = kr[hh * 4 + d4];
            const float* qq = q + hh * HD + 4 * d4;
            part[d4] = fmaf(qq[3], kk.w, fmaf(qq[2], kk.z, fmaf(qq[1], kk.y, qq[0] * kk.x)));
          }
          s[hh] = (part[0] + part[1]) + (part[2] + part[3]);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float mn = fmaxf(m[hh], s[hh]);
          const float c = exp2f(m[hh] - mn);   // 1 when the maximum stays, 0 for the first key (m = -inf)
          const float p = exp2f(s[hh] - mn);
          m[hh] = mn;
          l[hh] = fmaf(l[hh], c, p);
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 vv = vr[hh * 4 + d4];
            float* aa = acc + hh * HD + 4 * d4;
            aa[0] = fmaf(aa[0], c, p * vv.x);
            aa[1] = fmaf(aa[1], c, p * vv.y);
            aa[2] = fmaf(aa[2], c, p * vv.z);
            aa[3] = fmaf(aa[3], c, p * vv.w);
          }
        }
      }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const float inv = 1.0f / l[hh];
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[hh * HD + d] *= inv;
    }
  }
  __syncwarp();
  st_operand16(tl, ao_hi, ao_lo, half * 2 * HD, acc);
  st_operand16(tl, ao_hi, ao_lo, half * 2 * HD + HD, acc + HD);
}

// k | v columns of the in-projection (+ bias) into the shared token-major tile: the thread with half 0 moves k, half 1 v
__device__ __forceinline__ void kv_epilogue(uint32_t tl, float* __restrict__ kv, int row, int half, bool is_agent, int wtype,
                                            int col_agent, int col_task, const float* __restrict__ b_agent,
                                            const float* __restrict__ b_task) {
  const float* bias = (is_agent ? b_agent : b_task) + D + half * D;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    float v[16];
    ld_variant16(tl, col_agent + D + half * D, col_task + D + half * D, g * 16, is_agent, wtype, v);
    float4* dst = (float4*)&kv[row * KV_STRIDE + half * D + g * 16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = make_float4(v[4 * i] + __ldg(&bias[g * 16 + 4 * i]), v[4 * i + 1] + __ldg(&bias[g * 16 + 4 * i + 1]),
                           v[4 * i + 2] + __ldg(&bias[g * 16 + 4 * i + 2]), v[4 * i + 3] + __ldg(&bias[g * 16 + 4 * i + 3]));
  }
}

// x <- LayerNorm(x + acc + bias) (eps 1e-5, biased variance) on this thread's row; the two threads of a row own 32
// columns each and exchange partial sums through shared memory
// (returns the dot product of this thread's 32 output columns with dot_w, if given: the AttCommitNet heads)
__device__ __forceinline__ float ln_epilogue(uint32_t tl, int row, int half, int d_col, const float* __restrict__ bias,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             float (*s_part)[2][ROWS], float* __restrict__ tile = nullptr,
                                             const float* __restrict__ dot_w = nullptr) {
  float dot = 0.0f;
  float z[32];
  float sum = 0.0f;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int c = half * 32 + g * 16;
    float v[16], h[16], l[16];
    tmem_ld16(tl + d_col + c, v);
    tmem_ld16(tl + A0_HI + c, h);
    tmem_ld16(tl + A0_LO + c, l);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      z[g * 16 + i] = (h[i] + l[i]) + (v[i] + __ldg(&bias[c + i]));
      sum += z[g * 16 + i];
    }
  }
  s_part[0][half][row] = sum;
  worker_sync();
  const float mean = (s_part[0][0][row] + s_part[0][1][row]) * (1.0f / D);
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float d = z[i] - mean;
    var = fmaf(d, d, var);
  }
  s_part[1][half][row] = var;
  worker_sync();
  var = s_part[1][0][row] + s_part[1][1][row];
  const float inv = rsqrtf(var * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int c = half * 32 + g * 16;
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = (z[g * 16 + i] - mean) * inv * __ldg(&gamma[c + i]) + __ldg(&beta[c + i]);
    st_operand16(tl, A0_HI, A0_LO, c, y);
    if (tile) {   // a float32 copy of the row for the context pooling
      float4* dst = (float4*)&tile[row * ZG_STRIDE + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    }
    if (dot_w) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dot = fmaf(__ldg(&dot_w[c + i]), y[i], dot);
    }
  }
  return dot;
}

template <bool COMMIT>
__global__ void __launch_bounds__(NT, 1) att_tc_kernel(const __grid_constant__ Params P, int group) {
  const ChunkDesc* const chunks = chunk_table<COMMIT>();
  constexpr int N_CHUNKS = COMMIT ? NCHUNK_COMMIT : NCHUNK;            // streamed per pass
  constexpr int N_LAYER_CHUNKS = COMMIT ? NCHUNK_COMMIT : CH_PAIR1;    // consumed by the layer loop of the MMA thread
  constexpr int AFD = COMMIT ? 13 : AF;                                // agent features per token
  extern __shared__ __align__(1024) unsigned char dsm[];
  unsigned char* ring = dsm;
  float* kvg = (float*)(dsm + NSLOT * SLOT_BYTES);
  uint16_t* plist = (uint16_t*)(dsm + NSLOT * SLOT_BYTES + KVG_BYTES);
  __shared__ __align__(8) uint64_t s_full[NSLOT], s_empty[NSLOT], s_a_ready, s_d_ready;
  __shared__ uint32_t s_tmem;
  __shared__ int s_env[GLIST], s_na[GLIST], s_nt[GLIST], s_scan[NT / 32 + 1];
  __shared__ Seg s_seg[GMAX];
  __shared__ uint8_t s_seg_of[ROWS];
  __shared__ int s_nseg, s_R, s_split, s_next, s_nvalid, s_done;
  __shared__ float s_part[2][2][ROWS];
  __shared__ float s_ctx[GMAX][D], s_hc[GMAX][D];   // AttContextPairNet: context vector, its pair-head term
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = P.max_tasks, MA = P.max_agents;
  const float* w = P.w;
  const TcOffsets& o = P.o;

  // work split: CTA b of the persistent (one per SM) grid takes the counted environments of rank [b * per, (b + 1) * per)
  // and streams them through its passes (group > 0: `group` environments per CTA instead, A/B runs)
  const PickArgs PA{P.need, P.env_idx, P.n};
  const EnvScan ES = scan_envs(PA, s_scan);
  const int T = ES.total;
  const int per = group > 0 ? group : (T + (int)gridDim.x - 1) / (int)gridDim.x;
  const int stride = group > 0 ? (int)gridDim.x * per : T;   // group > 0: the grid strides over the groups
  if ((int)blockIdx.x * per >= T) return;
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 1);
    }
    mbar_init(&s_a_ready, NWORK);
    mbar_init(&s_d_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;

  uint32_t seq = 0;  // weight chunks consumed / produced so far (MMA and producer threads)
  uint32_t pa = 0;   // parity of s_a_ready (MMA thread)
  uint32_t pd = 0;   // parity of s_d_ready (workers)
  bool first_pass = true;
  int n_ts = 0;
  for (int r_lo = blockIdx.x * per; r_lo < T; r_lo += stride) {
  const int r_hi = min(T, r_lo + per);
  int cursor = r_lo;
  while (cursor < r_hi) {
  // the next (at most GLIST) environments of the range; a pass that the list leaves half empty is deferred to the next
  // list when the range holds more environments
  const int m = min(GLIST, r_hi - cursor);
  const bool more = cursor + m < r_hi;
  select_envs(PA, ES, cursor, m, s_env);
  __syncthreads();
  // valid rows / columns are a prefix by construction of the token builders: one warp per environment counts them
  for (int g = warp; g < m; g += NT / 32) {
    const int e = s_env[g];
    int na = 0, nt = 0;
    for (int base = 0; base < MA; base += 32) {
      const unsigned bal = __ballot_sync(0xffffffffu, base + lane < MA && P.agent_mask[(size_t)e * MA + base + lane] == 0);
      if (bal == 0xffffffffu) { na += 32; continue; }
      na += __ffs(~bal) - 1;
      break;
    }
    for (int base = 0; base < MT; base += 32) {
      const unsigned bal = __ballot_sync(0xffffffffu, base + lane < MT && P.task_mask[(size_t)e * MT + base + lane] == 0);
      if (bal == 0xffffffffu) { nt += 32; continue; }
      nt += __ffs(~bal) - 1;
      break;
    }
    if (lane == 0) {
      s_na[g] = na;
      s_nt[g] = nt;
    }
  }
  if (tid == 0) {
    s_next = 0;
    s_done = m;
  }
  // outputs of every listed environment start at zero (padded rows / columns, invalid edges)
  for (int g = 0; g < m; ++g) {
    if (COMMIT) {
      for (int idx = tid; idx < MT; idx += NT) P.pri[(size_t)s_env[g] * MT + idx] = 0.0f;
      for (int idx = tid; idx < MA; idx += NT) P.com[(size_t)s_env[g] * MA + idx] = 0.0f;
    } else {
      float* sc = P.scores + (size_t)s_env[g] * MA * MT;
      for (int idx = tid; idx < MA * MT; idx += NT) sc[idx] = 0.0f;
    }
  }
  __syncthreads();
  for (;;) {
    // ---- next pass: as many of the remaining environments as fit 128 rows (at most GMAX)
    if (tid == 0) {
      const int g_first = s_next;
      int g = s_next, ns = 0, sa = 0, st = 0;
      bool full = false;
      while (g < m && ns < GMAX) {
        const int na = s_na[g], nt = s_nt[g];
        if (COMMIT ? (na == 0 && nt == 0) : (na == 0 || nt == 0)) { ++g; continue; }
        if (((sa + na + 3) & ~3) + st + nt > ROWS) { full = true; break; }
        s_seg[ns].e = s_env[g];
        s_seg[ns].abase = sa;
        s_seg[ns].tbase = st;
        s_seg[ns].na = na;
        s_seg[ns].nt = nt;
        sa += na;
        st += nt;
        ++ns;
        ++g;
      }
      if (ns == GMAX) full = true;
      if (!full && more && ns > 0 && g_first > 0) {
        // the list ran out before the pass was full: take these environments up again with the next list
        s_done = g_first;
        ns = 0;
        sa = st = 0;
        g = m;
      }
      const int split = (sa + 3) & ~3;
      for (int q = 0; q < ns; ++q) s_seg[q].tbase += split;
      s_next = g;
      s_nseg = ns;
      s_split = split;
      s_R = split + st;
      s_nvalid = 0;
    }
    __syncthreads();
    const int nseg = s_nseg;
    if (nseg == 0) break;
    const int split = s_split;

    if (warp == 9) {
      // ================= producer: weight chunks through the ring
      if (lane == 0) {
        for (int c = 0; c < N_CHUNKS; ++c, ++seq) {
          const int slot = seq & (NSLOT - 1);
          const uint32_t use = seq / NSLOT;
          if (use > 0) mbar_wait_relaxed(&s_empty[slot], (use - 1) & 1);
          const ChunkDesc cd = chunks[c];
          const uint32_t bytes = (uint32_t)cd.nc * cd.kc * 8u;
          mbar_expect_tx(&s_full[slot], bytes);
          bulk_g2s(ring + slot * SLOT_BYTES, P.tcw + cd.off, bytes, &s_full[slot]);
        }
      }
      __syncwarp();
    } else if (warp == 8) {
      // ================= MMA issuer
      if (lane == 0) {
        int c = 0;
        while (c < N_LAYER_CHUNKS) {
          mbar_wait(&s_a_ready, pa);
          pa ^= 1;
          tc_fence_after();
          for (;;) {
            const ChunkDesc cd = chunks[c];
            const int slot = seq & (NSLOT - 1);
            mbar_wait(&s_full[slot], (seq / NSLOT) & 1);
            tc_fence_after();
            issue_chunk(tm, cd, ring + slot * SLOT_BYTES);
            tc_commit(&s_empty[slot]);
            ++seq;
            ++c;
            if (cd.last) break;
          }
          tc_commit(&s_d_ready);
        }
        if constexpr (!COMMIT) {
        // pair head: the two weight chunks stay in their slots for every tile of the pass
        mbar_wait(&s_a_ready, pa);
        pa ^= 1;
        const int nvalid = *(volatile int*)&s_nvalid;
        const int ntiles = (nvalid + ROWS - 1) / ROWS;
        const int slot1 = seq & (NSLOT - 1);
        mbar_wait(&s_full[slot1], (seq / NSLOT) & 1);
        ++seq;
        const int slot2 = seq & (NSLOT - 1);
        mbar_wait(&s_full[slot2], (seq / NSLOT) & 1);
        ++seq;
        ChunkDesc p1b = c_chunks[CH_PAIR1], p2b = c_chunks[CH_PAIR2];   // the second tile of a round
        p1b.a_hi = p2b.a_hi = U2_HI;
        p1b.a_lo = p2b.a_lo = U2_LO;
        p1b.d_col = P1B_D;
        p2b.d_col = P2B_D;
        for (int t = 0; t < ntiles; t += 2) {
          mbar_wait(&s_a_ready, pa);
          pa ^= 1;
          tc_fence_after();
          issue_chunk(tm, c_chunks[CH_PAIR1], ring + slot1 * SLOT_BYTES);
          if (t + 1 < ntiles) issue_chunk(tm, p1b, ring + slot1 * SLOT_BYTES);
          tc_commit(&s_d_ready);
          mbar_wait(&s_a_ready, pa);
          pa ^= 1;
          tc_fence_after();
          issue_chunk(tm, c_chunks[CH_PAIR2], ring + slot2 * SLOT_BYTES);
          if (t + 1 < ntiles) issue_chunk(tm, p2b, ring + slot2 * SLOT_BYTES);
          tc_commit(&s_d_ready);
        }
        tc_commit(&s_empty[slot1]);
        tc_commit(&s_empty[slot2]);
        }
      }
      __syncwarp();
    } else {
      // ================= workers
      const int q = warp & 3, half = warp >> 2;
      const int row = q * 32 + lane;
      const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);
      const int R = s_R;
      const bool is_agent = row < split;
      const int wtype = (q * 32 + 32 <= split) ? 0 : (q * 32 >= split ? 1 : 2);
      // the segment (environment) of this thread's row
      int sid = SEG_NONE;
      for (int q2 = 0; q2 < nseg; ++q2) {
        const Seg t = s_seg[q2];
        if ((row >= t.abase && row < t.abase + t.na) || (row >= t.tbase && row < t.tbase + t.nt)) sid = q2;
      }
      const bool on = sid != SEG_NONE;
      Seg sg = s_seg[on ? sid : 0];
      if (half == 0) s_seg_of[row] = (uint8_t)sid;   // read by the pair tiles (several worker barriers later)
      (void)R;

      // development: stage timestamps of CTA 0 / thread 0 behind the activation dumps
#if defined(MUAV_TC_DEBUG)
      long long* ts = (P.dbg && blockIdx.x == 0 && tid == 0) ? (long long*)(P.dbg + 4 * ROWS * D) : nullptr;
#define TS_MARK()                               \
  do {                                          \
    if (ts && n_ts < 250) ts[1 + n_ts++] = clock64(); \
  } while (0)
#else
      long long* const ts = nullptr;
#define TS_MARK() ((void)0)
#endif
      TS_MARK();
      // ---- raw features as the A operand of the projections (K padded to 16)
      if (half == 0) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = 0.0f;
        if (on) {
          if (is_agent) {
            const float* src = P.agent_feats + ((size_t)sg.e * MA + (row - sg.abase)) * AFD;
#pragma unroll
            for (int i = 0; i < AFD; ++i) f[i] = src[i];
          } else {
            const float* src = P.task_feats + ((size_t)sg.e * MT + (row - sg.tbase)) * TF;
#pragma unroll
            for (int i = 0; i < TF; ++i) f[i] = src[i];
          }
        }
        __syncwarp();
        st_operand16(tl, F_HI, F_LO, 0, f);
        tc_wait_st();
      }
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- list of the pairs with a valid edge (needs only the tokens: built while the first MMAs run)
      if constexpr (!COMMIT) {
        int poff[GMAX + 1];
        poff[0] = 0;
#pragma unroll
        for (int q2 = 0; q2 < GMAX; ++q2) poff[q2 + 1] = poff[q2] + (q2 < nseg ? s_seg[q2].na * s_seg[q2].nt : 0);
        const int np = poff[GMAX];
        for (int base = 0; base < np; base += NWORK) {
          const int idx = base + tid;
          bool v = false;
          int code = 0;
          if (idx < np) {
            int q2 = 0;
#pragma unroll
            for (int k = 1; k < GMAX; ++k)
              if (idx >= poff[k]) q2 = k;
            const Seg ps = s_seg[q2];
            const int loc = idx - poff[q2];
            const int i = loc / ps.nt, j = loc - i * ps.nt;
            v = P.edge_valid[(size_t)ps.e * MA * MT + (size_t)i * MT + j] != 0.0f;
            code = ((ps.abase + i) << 7) | (ps.tbase + j);
          }
          // one shared-memory atomic per warp
          const unsigned bal = __ballot_sync(0xffffffffu, v);
          if (bal) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&s_nvalid, __popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (v) plist[pos + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)code;
          }
        }
      }

      // ---- x = proj(feats) + type_embed
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      {
        const float* b = w + (is_agent ? o.agent_proj_b : o.task_proj_b);
        const float* te = w + o.type_embed + (is_agent ? 0 : D);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int c = half * 32 + g * 16;
          float v[16];
          ld_variant16(tl, 128, 192, c, is_agent, wtype, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] + __ldg(&b[c + i]) + __ldg(&te[c + i]);
          st_operand16(tl, A0_HI, A0_LO, c, v);
        }
        tc_wait_st();
        if (first_pass) dump_a0(P, 0, tl, row, half);
      }
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      if constexpr (COMMIT) {
        // ---- AttCommitNet: two encoder layers, then priority = sigmoid(w_p . h + b_p) on task rows and commit =
        // sigmoid(w_c . h + b_c) on agent rows (AttentionCommit.py:90-100)
        float dot = 0.0f;
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
          wait_d(&s_d_ready, pd, lane);
          kv_epilogue(tl, kvg, row, half, is_agent, 0, 128, 128, w + o.enc_in_b[l], w + o.enc_in_b[l]);
          worker_sync();
          attention(tl, kvg, half, on, sg, false, is_agent, 0, 128, 128, w + o.enc_in_b[l], w + o.enc_in_b[l], AO_HI, AO_LO);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&s_a_ready);
          wait_d(&s_d_ready, pd, lane);
          ln_epilogue(tl, row, half, 448, w + o.enc_out_b[l], w + o.enc_n1_w[l], w + o.enc_n1_b[l], s_part);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&s_a_ready);
          wait_d(&s_d_ready, pd, lane);
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            const int c = half * 64 + g * 16;
            float v[16];
            tmem_ld16(tl + 128 + c, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + __ldg(&w[o.enc_l1_b[l] + c + i]), 0.0f);
            st_operand16(tl, H_HI, H_LO, c, v);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&s_a_ready);
          wait_d(&s_d_ready, pd, lane);
          dot = ln_epilogue(tl, row, half, 128, w + o.enc_l2_b[l], w + o.enc_n2_w[l], w + o.enc_n2_b[l], s_part, nullptr,
                            l == 1 ? w + (is_agent ? o.commit_w : o.priority_w) : nullptr);
          tc_wait_st();
          tc_fence_before();
          if (l == 0) mbar_arrive(&s_a_ready);
        }
        s_part[0][half][row] = dot;
        worker_sync();
        if (half == 0 && on) {
          const float acc = s_part[0][0][row] + s_part[0][1][row] + w[is_agent ? o.commit_b : o.priority_b];
          const float v = 1.0f / (1.0f + expf(-acc));
          if (is_agent) P.com[(size_t)sg.e * MA + (row - sg.abase)] = v;
          else P.pri[(size_t)sg.e * MT + (row - sg.tbase)] = v;
        }
      } else {
      // ---- encoder self-attention
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      kv_epilogue(tl, kvg, row, half, is_agent, 0, 128, 128, w + o.enc_in_b[0], w + o.enc_in_b[0]);
      worker_sync();
      attention(tl, kvg, half, on, sg, false, is_agent, 0, 128, 128, w + o.enc_in_b[0], w + o.enc_in_b[0], AO_HI, AO_LO);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- x1 = LN1(x + out_proj(attn))
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      ln_epilogue(tl, row, half, 448, w + o.enc_out_b[0], w + o.enc_n1_w[0], w + o.enc_n1_b[0], s_part);
      tc_wait_st();
      if (first_pass) dump_a0(P, 1, tl, row, half);
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- hidden = relu(linear1(x1))
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int c = half * 64 + g * 16;
        float v[16];
        tmem_ld16(tl + 128 + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + __ldg(&w[o.enc_l1_b[0] + c + i]), 0.0f);
        st_operand16(tl, H_HI, H_LO, c, v);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- h = LN2(x1 + linear2(hidden))
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      ln_epilogue(tl, row, half, 128, w + o.enc_l2_b[0], w + o.enc_n2_w[0], w + o.enc_n2_b[0], s_part, o.has_context ? kvg : nullptr);
      tc_wait_st();
      if (first_pass) dump_a0(P, 2, tl, row, half);
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();
      if (o.has_context) {
        // ctx = ctx_proj(context) + mean of h over the environment's tokens (ContextPairHybrid.py:140-142), then its
        // term of the first pair-head layer hc = Wc ctx (head1 input rows 192..255); runs under the cross in-projection MMAs
        worker_sync();
        for (int idx = tid; idx < nseg * D; idx += NWORK) {
          const int g = idx / D, k = idx - g * D;
          const Seg ps = s_seg[g];
          float sum = 0.0f;
          for (int r = 0; r < ps.na; ++r) sum += kvg[(ps.abase + r) * ZG_STRIDE + k];
          for (int r = 0; r < ps.nt; ++r) sum += kvg[(ps.tbase + r) * ZG_STRIDE + k];
          float c = w[o.ctx_proj_b + k];
          const float* cx = P.context + (size_t)ps.e * 8;
#pragma unroll
          for (int q2 = 0; q2 < 8; ++q2) c = fmaf(cx[q2], w[o.ctx_proj_w + q2 * D + k], c);
          s_ctx[g][k] = c + sum / (float)(ps.na + ps.nt);
        }
        worker_sync();
        for (int idx = tid; idx < nseg * D; idx += NWORK) {
          const int g = idx / D, oo = idx - g * D;
          float hc = 0.0f;
          for (int k = 0; k < D; ++k) hc = fmaf(w[o.head1_w + (size_t)(3 * D + k) * D + oo], s_ctx[g][k], hc);
          s_hc[g][oo] = hc;
        }
      }

      // ---- cross attention: agent rows take q from cross_a2t (columns 128..) and serve as k / v of cross_t2a
      // (columns 320 + 64..); task rows the other way round
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      kv_epilogue(tl, kvg, row, half, is_agent, wtype, 320, 128, w + o.t2a_in_b, w + o.a2t_in_b);
      worker_sync();
      attention(tl, kvg, half, on, sg, true, is_agent, wtype, 128, 320, w + o.a2t_in_b, w + o.t2a_in_b, AO2_HI, AO2_LO);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- z = h + out_proj(ctx) (a' / t'): the A operand of the first pair-head layer and, in shared memory, the
      // source of the pair products
      float* zt = kvg;
      float* gt = kvg + ROWS * ZG_STRIDE;
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
      {
        const float* b = w + (is_agent ? o.a2t_out_b : o.t2a_out_b);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int c = half * 32 + g * 16;
          float v[16], h[16], l[16];
          ld_variant16(tl, 384, 448, c, is_agent, wtype, v);
          tmem_ld16(tl + A0_HI + c, h);
          tmem_ld16(tl + A0_LO + c, l);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (h[i] + l[i]) + (v[i] + __ldg(&b[c + i]));
          st_operand16(tl, A0_HI, A0_LO, c, v);
          float4* dst = (float4*)&zt[row * ZG_STRIDE + c];
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        tc_wait_st();
        if (first_pass) dump_a0(P, 3, tl, row, half);
      }
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- g = Wa a' (agent rows) / Wt t' + b1 (task rows); list of the pairs with a valid edge
      wait_d(&s_d_ready, pd, lane);
      TS_MARK();
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int c = half * 32 + g * 16;
        float v[16];
        ld_variant16(tl, 128, 192, c, is_agent, wtype, v);
        if (!is_agent) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(&w[o.head1_b + c + i]);
        }
        float4* dst = (float4*)&gt[row * ZG_STRIDE + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      worker_sync();
      const int nvalid = s_nvalid;
      tc_fence_before();
      mbar_arrive(&s_a_ready);
      TS_MARK();

      // ---- pair tiles: logits = w3 . relu(W2 relu(Wat (a' * t') + g_a + g_t) + b2) + b3, two tiles of 128 pairs per
      // round (the second one in the columns the encoder no longer needs): half the hand-offs with the MMA thread
      const int ntiles = (nvalid + ROWS - 1) / ROWS;
#pragma unroll 1
      for (int t = 0; t < ntiles; t += 2) {
        const int nt2 = min(2, ntiles - t);
        int ta[2], tt[2];
        bool valid[2];
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          const int p = (t + u2) * ROWS + row;
          valid[u2] = p < nvalid;
          const int pc = plist[valid[u2] ? p : 0];
          ta[u2] = pc >> 7;
          tt[u2] = pc & 127;
        }
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          if (u2 < nt2) {
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
              const int c = half * 32 + g * 16;
              const float4* za = (const float4*)&zt[ta[u2] * ZG_STRIDE + c];
              const float4* zb = (const float4*)&zt[tt[u2] * ZG_STRIDE + c];
              float u[16];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = za[i], b = zb[i];
                u[4 * i] = a.x * b.x;
                u[4 * i + 1] = a.y * b.y;
                u[4 * i + 2] = a.z * b.z;
                u[4 * i + 3] = a.w * b.w;
              }
              st_operand16(tl, u2 ? U2_HI : U_HI, u2 ? U2_LO : U_LO, c, u);
            }
          }
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&s_a_ready);
      TS_MARK();

        wait_d(&s_d_ready, pd, lane);
      TS_MARK();
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          if (u2 < nt2) {
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
              const int c = half * 32 + g * 16;
              const float4* ga = (const float4*)&gt[ta[u2] * ZG_STRIDE + c];
              const float4* gb = (const float4*)&gt[tt[u2] * ZG_STRIDE + c];
              float v[16];
              tmem_ld16(tl + (u2 ? P1B_D : P1_D) + c, v);
              if (o.has_context) {
                const float* hc = &s_hc[s_seg_of[ta[u2]]][c];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += hc[i];
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = ga[i], b = gb[i];
                v[4 * i] = fmaxf(v[4 * i] + a.x + b.x, 0.0f);
                v[4 * i + 1] = fmaxf(v[4 * i + 1] + a.y + b.y, 0.0f);
                v[4 * i + 2] = fmaxf(v[4 * i + 2] + a.z + b.z, 0.0f);
                v[4 * i + 3] = fmaxf(v[4 * i + 3] + a.w + b.w, 0.0f);
              }
              st_operand16(tl, u2 ? U2_HI : U_HI, u2 ? U2_LO : U_LO, c, v);
            }
          }
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&s_a_ready);
      TS_MARK();

        wait_d(&s_d_ready, pd, lane);
      TS_MARK();
        // the two threads of a row take one tile each
        if (half < nt2) {
          float logit = __ldg(&w[o.head3_b]);
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            float v[16];
            tmem_ld16(tl + (half ? P2B_D : P2_D) + g * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              logit = fmaf(__ldg(&w[o.head3_w + g * 16 + i]), fmaxf(v[i] + __ldg(&w[o.head2_b + g * 16 + i]), 0.0f), logit);
          }
          if (valid[half]) {
            const int a_row = ta[half], t_row = tt[half];
            const Seg ps = s_seg[s_seg_of[a_row]];
            const size_t off = (size_t)ps.e * MA * MT + (size_t)(a_row - ps.abase) * MT + (t_row - ps.tbase);
            P.scores[off] = tanhf(logit) * P.clamp * P.edge_valid[off];
          }
        }
        // the next round's products overwrite the operand columns: every read of this round is complete (wait::ld above)
        tc_fence_before();
      }
      }   // !COMMIT
      if (ts) ts[0] = n_ts;
    }
    first_pass = false;
    __syncthreads();  // the next pass reuses the segment table and every buffer
  }
  cursor += s_done;
  __syncthreads();
  }
  }
  // the two resident slots are released by commits that may still be in flight: wait for them before the CTA ends
  if (warp == 8 && lane == 0 && seq >= 2) {
    const uint32_t s1 = seq - 2, s2 = seq - 1;
    mbar_wait(&s_empty[s1 & (NSLOT - 1)], (s1 / NSLOT) & 1);
    mbar_wait(&s_empty[s2 & (NSLOT - 1)], (s2 / NSLOT) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

// ---- packing of the weights: every chunk as hi plane then lo plane, each [K/4][N][4] (K-major core matrices of 8 x 16 B)
struct PackSrc {
  // output feature n of chunk c comes from matrix 0 (n < nsplit) or matrix 1 (n - nsplit); src: float offset of W^T
  // ([in][out], as packed by the host for the fp32 kernel) in the parameter buffer, ldo its row stride
  int src[NCHUNK][2], ldo[NCHUNK][2], n0[NCHUNK][2], kreal[NCHUNK][2];
  int nsplit[NCHUNK], k0[NCHUNK];
};
template <bool COMMIT>
__global__ void tc_pack_kernel(const float* __restrict__ w, const __grid_constant__ PackSrc S, float* __restrict__ out) {
  const int c = blockIdx.y;
  const ChunkDesc cd = chunk_table<COMMIT>()[c];
  const int total = cd.nc * cd.kc;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx / cd.kc, k = idx - n * cd.kc;
    const int v = n < S.nsplit[c] ? 0 : 1;
    const int nn = v ? n - S.nsplit[c] : n;
    const int kk = S.k0[c] + k;
    float x = 0.0f;
    if (kk < S.kreal[c][v]) x = w[S.src[c][v] + (size_t)kk * S.ldo[c][v] + S.n0[c][v] + nn];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int pos = ((k >> 2) * cd.nc + n) * 4 + (k & 3);
    out[cd.off + pos] = hi;
    out[cd.off + total + pos] = x - hi;
  }
}

#if defined(MUAV_TC_DEBUG)
static float* g_dbg = nullptr;   // development builds only: where the next launch dumps its first pass
#endif

}  // namespace muav_tc

#if defined(MUAV_TC_DEBUG)
extern "C" void muav_tc_debug_buffer_(float* d_dbg) { muav_tc::g_dbg = d_dbg; }
#endif

extern "C" int64_t muav_att_pair_tc_floats(void) { return (int64_t)muav_tc::TCW_FLOATS; }

extern "C" int muav_att_pair_tc_pack(const float* d_params, const muav_attpair_offsets* offsets, float* d_tc_weights,
                                     void* stream) {
  using namespace muav_tc;
  if (!d_params || !offsets || !d_tc_weights) return -22;
  const muav_attpair_offsets& o = *offsets;
  PackSrc S;
  // one matrix: every output feature from (src, ldo, n0); K slice [k0, k0 + kc) of kreal input features
  auto one = [&](int c, int src, int ldo, int n0, int k0, int kreal) {
    S.src[c][0] = S.src[c][1] = src; S.ldo[c][0] = S.ldo[c][1] = ldo; S.n0[c][0] = S.n0[c][1] = n0;
    S.kreal[c][0] = S.kreal[c][1] = kreal; S.nsplit[c] = 1 << 20; S.k0[c] = k0;
  };
  auto two = [&](int c, int srcA, int kA, int srcB, int kB, int ldo, int k0) {   // [A; B] stacked along N, 64 rows each
    S.src[c][0] = srcA; S.src[c][1] = srcB; S.ldo[c][0] = S.ldo[c][1] = ldo; S.n0[c][0] = S.n0[c][1] = 0;
    S.kreal[c][0] = kA; S.kreal[c][1] = kB; S.nsplit[c] = D; S.k0[c] = k0;
  };
  two(0, o.agent_proj_w, AF, o.task_proj_w, TF, D, 0);
  for (int i = 0; i < 4; ++i) one(1 + i, o.enc_in_w, 3 * D, 0, 16 * i, D);
  one(5, o.enc_out_w, D, 0, 0, D);
  for (int i = 0; i < 2; ++i) one(6 + i, o.enc_l1_w, 2 * D, 0, 32 * i, D);
  for (int i = 0; i < 2; ++i) one(8 + i, o.enc_l2_w, D, 0, 64 * i, 2 * D);
  for (int i = 0; i < 4; ++i) one(10 + i, o.a2t_in_w, 3 * D, 0, 16 * i, D);
  for (int i = 0; i < 4; ++i) one(14 + i, o.t2a_in_w, 3 * D, 0, 16 * i, D);
  for (int i = 0; i < 2; ++i) two(18 + i, o.a2t_out_w, D, o.t2a_out_w, D, D, 32 * i);
  // pair_head.0: input features [0, 64) agent block, [64, 128) task block, [128, 192) product block
  for (int i = 0; i < 2; ++i) two(20 + i, o.head1_w, D, o.head1_w + D * D, D, D, 32 * i);
  one(22, o.head1_w + 2 * D * D, D, 0, 0, D);
  one(23, o.head2_w, D / 2, 0, 0, D);
  tc_pack_kernel<false><<<dim3(8, NCHUNK), 256, 0, (cudaStream_t)stream>>>(d_params, S, d_tc_weights);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

extern "C" int muav_att_pair_scores_tc(const float* d_params, const muav_attpair_offsets* offsets, const float* d_tc_weights,
                                       const float* d_task_feats, const uint8_t* d_task_mask, const float* d_agent_feats,
                                       const uint8_t* d_agent_mask, const float* d_edge_valid, const float* d_context,
                                       const int32_t* d_env_idx, const uint8_t* d_need, int n, int max_tasks, int max_agents,
                                       float score_clamp, float* d_scores, void* stream) {
  using namespace muav_tc;
  if (!d_params || !offsets || !d_tc_weights || !d_task_feats || !d_task_mask || !d_agent_feats || !d_agent_mask ||
      !d_edge_valid || !d_scores)
    return -22;
  if ((offsets->has_context != 0) != (d_context != nullptr)) return -22;
  if (n < 0 || max_tasks < 1 || max_agents < 1 || max_agents + max_tasks > 48 || max_agents > 16) return -22;
  if (n == 0) return 0;
  Params P;
  memset(&P, 0, sizeof(P));
  P.w = d_params;
  {
    const muav_attpair_offsets& a = *offsets;
    TcOffsets& o = P.o;
    o.agent_proj_w = a.agent_proj_w; o.agent_proj_b = a.agent_proj_b; o.task_proj_w = a.task_proj_w;
    o.task_proj_b = a.task_proj_b; o.type_embed = a.type_embed;
    o.enc_in_w[0] = a.enc_in_w; o.enc_in_b[0] = a.enc_in_b; o.enc_out_w[0] = a.enc_out_w; o.enc_out_b[0] = a.enc_out_b;
    o.enc_l1_w[0] = a.enc_l1_w; o.enc_l1_b[0] = a.enc_l1_b; o.enc_l2_w[0] = a.enc_l2_w; o.enc_l2_b[0] = a.enc_l2_b;
    o.enc_n1_w[0] = a.enc_n1_w; o.enc_n1_b[0] = a.enc_n1_b; o.enc_n2_w[0] = a.enc_n2_w; o.enc_n2_b[0] = a.enc_n2_b;
    o.a2t_in_w = a.a2t_in_w; o.a2t_in_b = a.a2t_in_b; o.a2t_out_w = a.a2t_out_w; o.a2t_out_b = a.a2t_out_b;
    o.t2a_in_w = a.t2a_in_w; o.t2a_in_b = a.t2a_in_b; o.t2a_out_w = a.t2a_out_w; o.t2a_out_b = a.t2a_out_b;
    o.head1_w = a.head1_w; o.head1_b = a.head1_b; o.head2_w = a.head2_w; o.head2_b = a.head2_b; o.head3_w = a.head3_w;
    o.head3_b = a.head3_b; o.ctx_proj_w = a.ctx_proj_w; o.ctx_proj_b = a.ctx_proj_b; o.has_context = a.has_context;
  }
  P.tcw = d_tc_weights;
  P.task_feats = d_task_feats;
  P.task_mask = d_task_mask;
  P.agent_feats = d_agent_feats;
  P.agent_mask = d_agent_mask;
  P.edge_valid = d_edge_valid;
  P.context = d_context;
  P.env_idx = d_env_idx;
  P.need = d_need;
  P.scores = d_scores;
  P.n = n;
  P.max_tasks = max_tasks;
  P.max_agents = max_agents;
  P.clamp = score_clamp;
#if defined(MUAV_TC_DEBUG)
  P.dbg = g_dbg;
#endif
  static bool set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(att_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM);
    if (e != cudaSuccess) return -1000 - (int)e;
    if (dev >= 0 && dev < 64) set[dev] = true;
  }
  // persistent grid, one CTA per SM (MUAV_SCORER_TC_GROUP = g > 0: g environments per CTA, grid-strided: A/B runs)
  static int n_sm[64];
  if (dev < 0 || dev >= 64 || n_sm[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev);
    if (v < 1) v = 1;
    if (dev >= 0 && dev < 64) n_sm[dev] = v;
  }
  const int sms = (dev >= 0 && dev < 64) ? n_sm[dev] : 148;
  int group = 0;
  const char* ge = getenv("MUAV_SCORER_TC_GROUP");
  if (ge) group = atoi(ge);
  if (group < 0) group = 0;
  int grid = n < sms ? n : sms;
  att_tc_kernel<false><<<grid, NT, DYN_SMEM, (cudaStream_t)stream>>>(P, group);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

extern "C" int64_t muav_att_commit_tc_floats(void) { return (int64_t)muav_tc::TCW_FLOATS_COMMIT; }

extern "C" int muav_att_commit_tc_pack(const float* d_params, const muav_attcommit_offsets* offsets, float* d_tc_weights,
                                       void* stream) {
  using namespace muav_tc;
  if (!d_params || !offsets || !d_tc_weights) return -22;
  const muav_attcommit_offsets& o = *offsets;
  PackSrc S;
  memset(&S, 0, sizeof(S));
  auto one = [&](int c, int src, int ldo, int n0, int k0, int kreal) {
    S.src[c][0] = S.src[c][1] = src; S.ldo[c][0] = S.ldo[c][1] = ldo; S.n0[c][0] = S.n0[c][1] = n0;
    S.kreal[c][0] = S.kreal[c][1] = kreal; S.nsplit[c] = 1 << 20; S.k0[c] = k0;
  };
  S.src[0][0] = o.agent_proj_w; S.src[0][1] = o.task_proj_w; S.ldo[0][0] = S.ldo[0][1] = D;
  S.kreal[0][0] = 13; S.kreal[0][1] = TF; S.nsplit[0] = D; S.k0[0] = 0;
  for (int l = 0; l < 2; ++l) {
    const int b = 1 + 9 * l;
    for (int i = 0; i < 4; ++i) one(b + i, o.enc_in_w[l], 3 * D, 0, 16 * i, D);
    one(b + 4, o.enc_out_w[l], D, 0, 0, D);
    for (int i = 0; i < 2; ++i) one(b + 5 + i, o.enc_l1_w[l], 2 * D, 0, 32 * i, D);
    for (int i = 0; i < 2; ++i) one(b + 7 + i, o.enc_l2_w[l], D, 0, 64 * i, 2 * D);
  }
  tc_pack_kernel<true><<<dim3(8, NCHUNK_COMMIT), 256, 0, (cudaStream_t)stream>>>(d_params, S, d_tc_weights);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

extern "C" int muav_att_commit_vectors_tc(const float* d_params, const muav_attcommit_offsets* offsets,
                                          const float* d_tc_weights, const float* d_task_feats, const uint8_t* d_task_mask,
                                          const float* d_agent_feats13, const uint8_t* d_agent_mask, const int32_t* d_env_idx,
                                          const uint8_t* d_need, int n, int max_tasks, int max_agents, float* d_priorities,
                                          float* d_commits, void* stream) {
  using namespace muav_tc;
  if (!d_params || !offsets || !d_tc_weights || !d_task_feats || !d_task_mask || !d_agent_feats13 || !d_agent_mask ||
      !d_priorities || !d_commits)
    return -22;
  if (n < 0 || max_tasks < 1 || max_agents < 1 || max_agents + max_tasks > 48 || max_agents > 16) return -22;
  if (n == 0) return 0;
  Params P;
  memset(&P, 0, sizeof(P));
  P.w = d_params;
  {
    const muav_attcommit_offsets& a = *offsets;
    TcOffsets& o = P.o;
    o.agent_proj_w = a.agent_proj_w; o.agent_proj_b = a.agent_proj_b; o.task_proj_w = a.task_proj_w;
    o.task_proj_b = a.task_proj_b; o.type_embed = a.type_embed;
    for (int l = 0; l < 2; ++l) {
      o.enc_in_w[l] = a.enc_in_w[l]; o.enc_in_b[l] = a.enc_in_b[l]; o.enc_out_w[l] = a.enc_out_w[l];
      o.enc_out_b[l] = a.enc_out_b[l]; o.enc_l1_w[l] = a.enc_l1_w[l]; o.enc_l1_b[l] = a.enc_l1_b[l];
      o.enc_l2_w[l] = a.enc_l2_w[l]; o.enc_l2_b[l] = a.enc_l2_b[l]; o.enc_n1_w[l] = a.enc_n1_w[l];
      o.enc_n1_b[l] = a.enc_n1_b[l]; o.enc_n2_w[l] = a.enc_n2_w[l]; o.enc_n2_b[l] = a.enc_n2_b[l];
    }
    o.priority_w = a.priority_w; o.priority_b = a.priority_b; o.commit_w = a.commit_w; o.commit_b = a.commit_b;
  }
  P.tcw = d_tc_weights;
  P.task_feats = d_task_feats;
  P.task_mask = d_task_mask;
  P.agent_feats = d_agent_feats13;
  P.agent_mask = d_agent_mask;
  P.env_idx = d_env_idx;
  P.need = d_need;
  P.pri = d_priorities;
  P.com = d_commits;
  P.n = n;
  P.max_tasks = max_tasks;
  P.max_agents = max_agents;
  static bool set[64];
  static int n_sm[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(att_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM);
    if (e != cudaSuccess) return -1000 - (int)e;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev);
    if (dev >= 0 && dev < 64) {
      n_sm[dev] = v < 1 ? 1 : v;
      set[dev] = true;
    }
  }
  const int sms = (dev >= 0 && dev < 64 && n_sm[dev] > 0) ? n_sm[dev] : 148;
  att_tc_kernel<true><<<n < sms ? n : sms, NT, DYN_SMEM, (cudaStream_t)stream>>>(P, 0);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}
