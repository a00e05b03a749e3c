// gemm.cu -- persistent tcgen05 / TMEM / TMA GEMM for the dense layers of the path: NatureConvBody convolutions as
// shifted-row implicit GEMMs, fc4, heads (network_bodies.py:27-33, network_heads.py:18-21), forward, dgrad and wgrad.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T        bf16 operands, fp32 accumulation in tensor memory
//
// Operand storage is described per operand:
//   K-major  (major = 0): row-major [rows = M or N][K], K contiguous            (activations x weights: y = x W^T)
//   MN-major (major = 1): row-major [K][rows = M or N], M/N contiguous          (weight gradients: dW = g^T x)
// so no operand is ever transposed in memory.
//
// Convolutions without im2col ("tap addressing").  Activations live as [batch * G * G][C] matrices over a G x G grid
// per image.  A k x k / stride-1 convolution over that grid is  y[r] = sum_taps x[r + dy*G + dx] W_tap^T : the A operand
// of k-tile `kt` is the SAME matrix read at a row offset that depends on the tap the k-tile belongs to (rows that run
// off the grid produce garbage output rows, which the epilogue drops or the next layer never reads).  Stride-s
// convolutions are first turned into stride-1 ones by a space-to-depth(s) layout of their input, which the PRODUCING
// kernel writes directly (the replay gather for conv1, conv1's epilogue for conv2).  dgrad is the same with negative
// shifts; wgrad reads both operands MN-major with the tap shift applied to the B operand per output column block.
//
// One persistent CTA per SM loops over output tiles (128 x BN):
//   warp 0    TMA producer   cp.async.bulk.tensor into a STAGES-deep 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1    MMA issuer     one ELECTED thread (elect.sync, see elect_one): tcgen05.mma.cta_group::1.kind::f16,
//                            accumulators double-buffered in TMEM
//   warps 2-5 epilogue of even tiles, warps 6-9 epilogue of odd tiles (one group per accumulator stage):
//                            tcgen05.ld (32 lanes x 32 columns per warp) -> bias / ReLU -> bf16 | fp32 | atomic fp32,
//                            overlapped with the next tile's MMAs; warp 2 also owns TMEM alloc / dealloc
// Every kernel runs its prologue (barrier init, tensor-map prefetch, TMEM allocation) before pdl_sync(): under
// programmatic dependent launch that part overlaps the tail of the previous kernel (common.cuh).
// sm_100a only.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace b2rl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;           // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int GEMM_THREADS = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5 / 6-9: epilogue of even / odd tiles

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(bar)) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase completes instead of polling
// the barrier word -- eight epilogue warps spinning on shared memory measurably slow the tensor pipe's operand reads
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(s2u(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s2u(smem_dst)),
      "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// One elected lane of a fully active warp.  Code guarded by this predicate (instead of `lane == 0`) lets the compiler issue
// the uniform-datapath instructions (UTCHMMA, UTMALDG, UTCBAR) directly; behind a plain lane test it wraps every one
// of them in an ELECT / BRA.U.ANY loop, which costs ~50 extra cycles per MMA on the issuing thread.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptor halves: the single MMA-issuing thread must stay lean (every instruction it executes is on the critical
// path of the tensor pipe), so the 64-bit descriptors are kept as a constant high word and a low word that only
// receives 32-bit adds.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128B swizzle, version 1 (Blackwell)
//   K-major : 8-row groups 1024 B apart (SBO), LBO = 1 (x16 B)
//   MN-major: 64-element MN groups `lbo` bytes apart, 8-K-row groups 1024 B apart (SBO)
// base_offset (bits 49-51) = (start address >> 7) & 7 when the start is not aligned to the 1024-byte swizzle pattern
// (a window that begins s rows into a swizzled slab); 0 for pattern-aligned tiles.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;       // version
  d |= (uint64_t)2 << 61;       // SWIZZLE_128B
  return d;
}

struct GemmParams {
  int M, N, K;
  int ldd;                 // row stride of D in elements
  int k_tiles_per_split;   // K tiles (of 64) handled by one blockIdx.z
  int a_mn, b_mn;          // operand majors (0 = K-major, 1 = MN-major)
  int relu, out_mode;      // out_mode 0: bf16 store, 1: fp32 store, 2: fp32 atomicAdd (split-K / accumulate)
  const float* bias;       // [N] or null (added by split 0 only)
  void* D;
  // tap addressing (0 = plain GEMM)
  int a_tap_tiles;         // K-major A: k-tiles per tap (= channels / 64); the A matrix is [rows][a_tap_tiles*64]
  int b_tap_tiles;         // MN-major B (wgrad): output column tiles per tap (= channels / BN); B is [rows][b_tap_tiles*BN]
  int taps_x, grid_w, shift_sign;   // row shift of tap t = sign * ((t / taps_x) * grid_w + t % taps_x)
  // output row mapping (rows of the GEMM are positions of a G x G grid per image)
  int out_map;             // 0: identity; 1: G-grid -> space-to-depth(2) rows, valid V x V; 2: G-grid -> compact V x V
  int G, V;
  // dual launch: the same GEMM for two independent operand sets (online / target network) in ONE grid -- the first half of
  // the CTAs works on (tmA, tmB, D, bias), the second half on (tmA2, tmB2, D2, bias2); halves the per-kernel fixed cost
  int dual;
  void* D2;
  const float* bias2;
  // backward extras (see epilogue_tile): ReLU-gradient mask, bias-gradient accumulation, grid scatter maps 3 / 4
  const __nv_bfloat16* mask;
  int64_t mask_ld;
  float* dbias;
  int dbias_mod, sub_c;
  // out_mode 3: split-K with an in-kernel fix-up -- every split stores its fp32 partial tile to ws[split][row][col], the split
  // that arrives last at counters[tile] adds the partials in split order (deterministic), applies bias / ReLU and stores bf16
  float* ws;
  int ws_rows, ws_ld;
  int* counters;
};

__device__ __forceinline__ int tap_shift(const GemmParams& p, int tap) {
  return p.shift_sign * ((tap / p.taps_x) * p.grid_w + tap % p.taps_x);
}

// Epilogue of one 128 x BN tile for one warp (TMEM lane quarter q): wait for the accumulator stage, read it 32 columns
// at a time, hand the stage back to the MMA warp, apply bias / ReLU and store through the output row map.
// The accumulator of a tile may be kept as ACC partial sums (the MMA issuer rotates over them, the epilogue adds them).
// Optional in-kernel timeline (scripts/microbench/slab_trace.py builds a private copy of the library with -DB2RL_TRACE):
// clock64 stamps per CTA / role / tile, never compiled into the shipped library.
#ifdef B2RL_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;
#define B2RL_TRACE_AT(role, it, slot)                                                                            \
  do {                                                                                                           \
    if (g_trace_buf && (it) < 16)                                                                                \
      g_trace_buf[((((size_t)blockIdx.x * 4 + (role)) * 16 + (it)) * 4) + (slot)] = (unsigned long long)clock64(); \
  } while (0)
#else
#define B2RL_TRACE_AT(role, it, slot)
#endif

// scripts/microbench/umma_rate.cu measured on B200: one issuing thread sustains one tcgen05.mma (M=128, K=16) every
// max(53, N/2) cycles whatever the operand majors and whether consecutive MMAs hit the same accumulator or not, so
// ACC = 1 is used; small-N tiles are bound by that 53-cycle issue floor, not by accumulator dependencies.
// Backward extras (dgrad GEMMs): `mask` is the saved forward activation in the GEMM's own output coordinates -- the ReLU
// gradient is applied in the epilogue (v = mask > 0 ? v : 0); `dbias` receives the column sums of the masked tile (the bias
// gradient of the layer below) through a per-CTA shared accumulator `s_dbias`, index = column % dbias_mod.  Output row maps
// 3 / 4 scatter the tile into the G x G grid matrix the next backward GEMMs read:
//   map 3: rows are space-to-depth(2) positions [b][oy][ox] of an (V/2)^2 grid, columns are 4 sub-positions x sub_c channels
//          -> grid row b*G*G + (2*oy + dy)*G + 2*ox + dx, column = channel           (conv2's input gradient -> conv1's grid)
//   map 4: rows are images b, columns are V*V positions x sub_c channels
//          -> grid row b*G*G + (pos / V)*G + pos % V, column = channel               (fc4's input gradient -> conv3's grid)
// Grid rows that no tile covers keep whatever the destination holds: the caller keeps it zeroed (persistent buffer).
template <int BN, int ACC, bool EXT>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int m0, int n0, int q, int lane, uint32_t tmem_base,
                                              uint32_t as, uint32_t parity, bool has_acc, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, float* s_dbias, uint32_t trace_it = 1u << 30) {
  const bool tr = q == 0 && lane == 0;
  if (tr) B2RL_TRACE_AT(2, trace_it, 0);
  const int row = m0 + q * 32 + lane;
  bool valid = row < p.M;
  int64_t drow = row;
  int dcol0 = 0;
  int img = 0, sy = 0, sx = 0;                        // maps 3 / 4: image and source position of this row
  if ((p.out_map == 1 || p.out_map == 2) && valid) {
    const int gg = p.G * p.G;
    const int b = row / gg, rem = row - b * gg;
    const int oy = rem / p.G, ox = rem - oy * p.G;
    valid = oy < p.V && ox < p.V;
    if (p.out_map == 1) {
      const int h = p.V >> 1;
      drow = (int64_t)b * h * h + (oy >> 1) * h + (ox >> 1);
      dcol0 = (((oy & 1) << 1) | (ox & 1)) * p.N;
    } else {
      drow = (int64_t)b * p.V * p.V + oy * p.V + ox;
    }
  } else if (EXT && p.out_map == 3 && valid) {
    const int h = p.V >> 1, hh = h * h;
    img = row / hh;
    const int rem = row - img * hh;
    sy = rem / h, sx = rem - sy * h;
  } else if (EXT && p.out_map == 4) {
    img = row;
  }
  // backward extras: the ReLU mask of this thread's row (the saved forward activation) does not depend on the accumulator --
  // all of it is requested BEFORE the wait for the MMAs, so that its latency is hidden behind them
  int4 mk[EXT ? BN / 8 : 1];
  if constexpr (EXT) {
    if (p.mask && valid) {
      const int4* mp = reinterpret_cast<const int4*>(p.mask + (int64_t)row * p.mask_ld + n0);
#pragma unroll
      for (int j = 0; j < BN / 8; ++j)
        if (n0 + 8 * j < p.N) mk[j] = __ldg(mp + j);
    }
  }
  if (has_acc) {
    mb_wait(&tmem_full[as], parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tr) B2RL_TRACE_AT(2, trace_it, 1);
#pragma unroll
  for (int c = 0; c < BN; c += 32) {
    uint32_t r[32];
    if (has_acc) {
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + as * (ACC * BN) + c;
      tmem_ld32(t0, r);
#pragma unroll
      for (int a = 1; a < ACC; ++a) {
        uint32_t r2[32];
        tmem_ld32(t0 + a * BN, r2);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0;
    }
    if (c + 32 >= BN && has_acc) {
      // all TMEM reads of this warp for this tile are done: hand the accumulator stage back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mb_arrive(&tmem_empty[as]);
      if (tr) B2RL_TRACE_AT(2, trace_it, 2);
    }
    const bool live = valid && n0 + c < p.N;
    if (live) {
      if (p.bias && blockIdx.z == 0) {
        const float* bp = p.bias + n0 + c;
        if (n0 + c + 32 <= p.N && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {                             // 8 broadcast 16-byte loads instead of 32 scalar ones
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
            r[j] = __float_as_uint(__uint_as_float(r[j]) + b4.x);
            r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + b4.y);
            r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + b4.z);
            r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + b4.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < p.N) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(bp + j));
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.0f));
      }
      if constexpr (EXT) {
        if (p.mask) {                                                   // ReLU gradient: zero where the forward output was <= 0
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int4 m4 = mk[(c >> 3) + (j >> 3)];
            const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&m4);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 mf = __bfloat1622float2(mh[t]);
              if (!(mf.x > 0.0f)) r[j + 2 * t] = 0;
              if (!(mf.y > 0.0f)) r[j + 2 * t + 1] = 0;
            }
          }
        }
      }
    }
    if (EXT && p.dbias) {
      // column sums over the 32 rows of this warp: 5-step register transpose-reduce (31 shuffles), lane l ends up with the
      // total of column l; rows that are not live contribute zeros.  All 32 lanes take part.
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = live ? __uint_as_float(r[j]) : 0.0f;
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool hi = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
          const float send = hi ? v[j] : v[j + s];
          const float keep = hi ? v[j + s] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      if (n0 + c + lane < p.N) {
        if (p.dbias_mod > 0) atomicAdd(s_dbias + (n0 + c + lane) % p.dbias_mod, v[0]);
        else atomicAdd(p.dbias + n0 + c + lane, v[0]);             // dbias_mod 0: one bias per output column, straight to global
      }
    }
    if (live) {
      int64_t off;
      if (EXT && p.out_map == 3) {
        const int sub = (n0 + c) / p.sub_c, cc = (n0 + c) - sub * p.sub_c;
        off = ((int64_t)img * p.G * p.G + (2 * sy + (sub >> 1)) * p.G + 2 * sx + (sub & 1)) * p.ldd + cc;
      } else if (EXT && p.out_map == 4) {
        const int pos = (n0 + c) / p.sub_c, cc = (n0 + c) - pos * p.sub_c;
        off = ((int64_t)img * p.G * p.G + (pos / p.V) * p.G + pos % p.V) * p.ldd + cc;
      } else {
        off = drow * p.ldd + dcol0 + n0 + c;
      }
      if (p.out_mode == 0) {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + off;
        if (n0 + c + 32 <= p.N && (off % 8 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            int4 v;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              h[t] = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * t]), __uint_as_float(r[j + 2 * t + 1]));
            *reinterpret_cast<int4*>(d + j) = v;
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < p.N) d[j] = __float2bfloat16_rn(__uint_as_float(r[j]));
        }
      } else if (p.out_mode == 1) {
        float* d = reinterpret_cast<float*>(p.D) + off;
        if (n0 + c + 32 <= p.N && (off % 4 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        } else {
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < p.N) d[j] = __uint_as_float(r[j]);
        }
      } else {
        float* d = reinterpret_cast<float*>(p.D) + off;
        if (n0 + c + 32 <= p.N && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(d + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        } else {
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < p.N) atomicAdd(d + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  if (tr) B2RL_TRACE_AT(2, trace_it, 3);
}

// Epilogue of a split-K GEMM with in-kernel fix-up (out_mode 3; fc4 forward at batch 512: 32 output tiles cannot fill 148 SMs,
// and the 196 dependent MMAs of one un-split tile take ~5 us at the 53-cycle issue floor).  Replaces zero-fill + atomic
// split-K + bias/ReLU pass (three launches) by one launch.  Every CTA owns exactly ONE tile (the host sizes the grid so).
// Warp group 0 (warps 2-5) drains the accumulator into the fp32 scratch; then all EIGHT epilogue warps of the CTA that arrived
// last at the tile's counter add the `splits` partials -- each thread requests all its partials of a 32-column chunk before
// adding them (in split order: deterministic), so the fix-up costs about one L2 round trip.
template <int BN>
__device__ __forceinline__ void epilogue_fixup(const GemmParams& p, int tile, int m0, int n0, int warp, int lane,
                                               uint32_t tmem_base, bool has_acc, uint64_t* tmem_full, uint64_t* tmem_empty,
                                               int* s_flag) {
  const int q = warp & 3, grp = (warp - 2) >> 2;
  const int row = m0 + q * 32 + lane;
  const int splits = (int)gridDim.z;
  if (grp == 0) {
    if (has_acc) {
      mb_wait(&tmem_full[0], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    float* wrow = p.ws + ((int64_t)blockIdx.z * p.ws_rows + row) * p.ws_ld + n0;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      if (has_acc) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        __stcg(reinterpret_cast<float4*>(wrow + c + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
    }
    if (has_acc) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mb_arrive(&tmem_empty[0]);
    }
    __threadfence();
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (warp == 2 && lane == 0) *s_flag = atomicAdd(p.counters + tile, 1) == splits - 1;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (!*s_flag) return;
  __threadfence();
  if (warp == 2 && lane == 0) p.counters[tile] = 0;                       // re-armed for the next launch
  const bool valid = row < p.M;
  for (int c = grp * 32; c < BN; c += 64) {                                // the two warp groups take alternate 32-column chunks
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    for (int sp0 = 0; sp0 < splits; sp0 += 4) {                            // 4 splits x 8 float4 requested before any is added
      float4 v[4][8];
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        if (sp0 + s4 < splits) {
          const float4* src = reinterpret_cast<const float4*>(p.ws + ((int64_t)(sp0 + s4) * p.ws_rows + row) * p.ws_ld + n0 + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[s4][j] = __ldcg(src + j);
        }
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        if (sp0 + s4 < splits) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[4 * j] += v[s4][j].x; acc[4 * j + 1] += v[s4][j].y; acc[4 * j + 2] += v[s4][j].z; acc[4 * j + 3] += v[s4][j].w;
          }
        }
      }
    }
    if (valid && n0 + c < p.N) {
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c + j < p.N) acc[j] += __ldg(p.bias + n0 + c + j);
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.0f);
      }
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + (int64_t)row * p.ldd + n0 + c;
      if (n0 + c + 32 <= p.N && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          int4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(acc[j + 2 * t], acc[j + 2 * t + 1]);
          *reinterpret_cast<int4*>(d + j) = o;
        }
      } else {
        for (int j = 0; j < 32; ++j)
          if (n0 + c + j < p.N) d[j] = __float2bfloat16_rn(acc[j]);
      }
    }
  }
}

template <int BN, int STAGES, bool EXT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmB,
                                                                       const __grid_constant__ CUtensorMap tmA2,
                                                                       const __grid_constant__ CUtensorMap tmB2,
                                                                       const GemmParams p0) {
  const int n_cta = p0.dual ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool second = p0.dual && (int)blockIdx.x >= n_cta;
  const int cta = second ? (int)blockIdx.x - n_cta : (int)blockIdx.x;
  const CUtensorMap* mA = second ? &tmA2 : &tmA;
  const CUtensorMap* mB = second ? &tmB2 : &tmB;
  GemmParams p = p0;
  if (second) { p.D = p0.D2; p.bias = p0.bias2; }
  constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2, B_BYTES = BN * GEMM_BK * 2;
  constexpr int ACC = 1;     // partial accumulators per tile: measured, independent chains do not raise the MMA rate
  constexpr uint32_t TMEM_COLS = 2 * ACC * BN < 32 ? 32 : 2 * ACC * BN;   // two accumulator stages
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int kt_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kt_begin = blockIdx.z * p.k_tiles_per_split;
  const int kt_end = min(kt_total, kt_begin + p.k_tiles_per_split);
  const int n_kt = max(kt_end - kt_begin, 0);

  __shared__ float s_dbias[128];                    // per-CTA bias-gradient accumulator (backward extras)
  __shared__ int s_fix[2];                          // out_mode 3: "this CTA arrived last at the tile's counter"
  if (threadIdx.x < 128) s_dbias[threadIdx.x] = 0.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mb_init(&tmem_full[s], 1); mb_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // everything above (barriers, tensor-map prefetch, TMEM allocation) overlaps the previous kernel's tail

  if (warp == 0 && n_kt > 0 && elect_one()) {
    // ---------------------------------------------------------------------- TMA producer
    uint32_t it = 0;                                         // global k-iteration counter across tiles
    for (int tile = cta; tile < tiles; tile += n_cta) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int m0 = mt * GEMM_BM, n0 = nt * BN;
      int b_shift = 0, b_col = n0;
      if (p.b_tap_tiles > 0) {                               // wgrad: the tap is selected by the output column block
        const int tap = nt / p.b_tap_tiles;
        b_shift = tap_shift(p, tap);
        b_col = (nt - tap * p.b_tap_tiles) * BN;
      }
      for (int i = 0; i < n_kt; ++i, ++it) {
        const int s = it % STAGES;
        mb_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mb_expect_tx(&full[s], A_BYTES + B_BYTES);
        const int g = kt_begin + i;
        uint8_t* a = sA + s * A_BYTES;
        uint8_t* b = sB + s * B_BYTES;
        if (!p.a_mn) {
          int ak = g * GEMM_BK, arow = m0;
          if (p.a_tap_tiles > 0) {
            const int tap = g / p.a_tap_tiles;
            ak = (g - tap * p.a_tap_tiles) * GEMM_BK;
            arow = m0 + tap_shift(p, tap);
          }
          tma_load_2d(a, mA, &full[s], ak, arow);                       // box [128 rows][64 k]
        } else {
          tma_load_2d(a, mA, &full[s], m0, g * GEMM_BK);                // 2 boxes [64 k][64 m]
          tma_load_2d(a + 8192, mA, &full[s], m0 + 64, g * GEMM_BK);
        }
        if (!p.b_mn) {
          tma_load_2d(b, mB, &full[s], g * GEMM_BK, n0);                // box [BN rows][64 k]
        } else {
#pragma unroll
          for (int q = 0; q < (BN + 63) / 64; ++q)
            tma_load_2d(b + q * 8192, mB, &full[s], b_col + q * 64, g * GEMM_BK + b_shift);
        }
      }
    }
  } else if (warp == 1 && n_kt > 0 && elect_one()) {
    // ---------------------------------------------------------------------- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
    const uint32_t a_lo0 = desc_lo(s2u(sA), p.a_mn ? 8192 : 16), b_lo0 = desc_lo(s2u(sB), p.b_mn ? 8192 : 16);
    const uint32_t a_step = p.a_mn ? 128 : 2, b_step = p.b_mn ? 128 : 2;
    uint32_t it = 0, tcount = 0;
    for (int tile = cta; tile < tiles; tile += n_cta, ++tcount) {
      const uint32_t as = tcount & 1;
      mb_wait(&tmem_empty[as], ((tcount >> 1) & 1) ^ 1);     // epilogue drained this accumulator stage
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + as * (ACC * BN);
      for (int i = 0; i < n_kt; ++i, ++it) {
        const int s = it % STAGES;
        mb_wait(&full[s], (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // K-major: 16 k-elements = 32 bytes along the swizzled row; MN-major: 16 k-rows = 2048 bytes (in 16-byte units: 2 / 128)
        uint32_t a_lo = a_lo0 + (uint32_t)s * (A_BYTES >> 4), b_lo = b_lo0 + (uint32_t)s * (B_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // k16 step k of every k-tile goes to partial accumulator k % ACC: ACC independent accumulation chains
          umma_f16_lh(acc + (k % ACC) * BN, a_lo, b_lo, idesc, (i > 0 || k >= ACC) ? 1u : 0u);
          a_lo += a_step;
          b_lo += b_step;
        }
        umma_commit(&empty[s]);                    // frees the smem stage when these MMAs retire
      }
      umma_commit(&tmem_full[as]);                 // accumulator of this tile complete
    }
  } else if (warp >= 2) {
    // ---------------------------------------------------------------------- epilogue (warp%4 selects the TMEM lane quarter)
    if (!EXT && p.out_mode == 3) {
      // split-K with in-kernel fix-up: one tile per CTA, all eight epilogue warps take part in the fix-up
      if (cta < tiles) {
        const int mt = cta / n_tiles, nt = cta - mt * n_tiles;
        epilogue_fixup<BN>(p, cta, mt * GEMM_BM, nt * BN, warp, lane, tmem_base, n_kt > 0, tmem_full, tmem_empty, s_fix);
      }
    } else {
      // two groups of four warps, one per accumulator stage: the epilogue of tile t overlaps that of tile t+1
      const int q = warp & 3;
      const uint32_t grp = (uint32_t)(warp - 2) >> 2;
      uint32_t tcount = 0;
      for (int tile = cta; tile < tiles; tile += n_cta, ++tcount) {
        if ((tcount & 1) != grp) continue;
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int m0 = mt * GEMM_BM, n0 = nt * BN;
        epilogue_tile<BN, ACC, EXT>(p, m0, n0, q, lane, tmem_base, tcount & 1, (tcount >> 1) & 1, n_kt > 0, tmem_full, tmem_empty, s_dbias);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (EXT && p.dbias && p.dbias_mod > 0 && (int)threadIdx.x < p.dbias_mod) atomicAdd(p.dbias + threadIdx.x, s_dbias[threadIdx.x]);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Slab variant of the forward / dgrad convolution GEMM: the taps of a tile read overlapping row windows of the same
// activation matrix, so the CTA loads ONE slab of 128 + max_shift rows per tile and issues every tap's MMAs on windows
// that start `shift` rows (x 128 bytes) into that swizzled slab.  The 128-byte swizzle is a pure function of the shared
// memory ADDRESS (bits 4-6 ^= bits 7-9), for TMA writes and UMMA reads alike, so a window may start at any row of a
// 1024-byte-aligned slab with descriptor base_offset = 0 (measured on B200: base_offset = (start >> 7) & 7 gives wrong
// results, 0 gives the right ones).  The weights (all taps) are loaded once per CTA and stay resident.  Activation traffic drops by the
// number of taps (4x conv1/conv2, 9x conv3) and the weight traffic per tile to zero.
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// K1: fused replay gather -> exact u8->bf16 -> conv1 operand.  Instead of reading a materialised bf16 space-to-depth matrix
// (57.8 MB per batch-512 stack, written by the gather kernel and re-read here), conv1's forward and weight-gradient kernels
// build their activation slabs themselves from the uint8 frame ring: four converter warps read, for every slab row
// (b, gy, gx) and 16-byte chunk j (8 channels = frame j/2, pixel rows 4gy + 2(j&1) + {0,1}, pixel columns 4gx..4gx+3), two
// aligned 32-bit words of the ring, convert the eight pixels exactly (integers 0..255 are representable in bf16; the 1/255
// of ImageNormalizer is folded into conv1's weights) and store one 16-byte chunk at the 128-byte-swizzled position
// (chunk ^ (row & 7)) the TMA would have written.  Reference chain replaced: replay.py:124-134 (frame-stack gather),
// normalizer.py:58-61, network_bodies.py:27.
// ---------------------------------------------------------------------------------------------------------------
struct U8Src {
  const uint8_t* frames;     // ring [capacity][row_bytes]
  const int64_t* idx;        // sampled ring indices [batch]
  int64_t row_bytes;         // bytes per ring row (84 * 84)
  int first;                 // ring row of channel-frame 0 relative to idx[b]: -(history-1) for s, n_step-(history-1) for s'
  int frame_w;               // pixels per frame row (84)
  int G;                     // grid width = frame_w / 4 (21); slab rows are (b, gy, gx) over G x G positions per image
  int rows;                  // batch * G * G
  int stages;                // uint8 staging tiles in flight (set by the launcher, <= U8_MAX_STAGES)
};

__device__ __forceinline__ int4 cvt8_u8_bf16(uint32_t w0, uint32_t w1) {
  // exact u8 -> bf16 without I2F: 0x4B0000vv is the float 2^23 + v; subtract 2^23; the top 16 bits are the bf16
  const float m = 8388608.0f;
  uint32_t f[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[k] = __float_as_uint(__uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7540 + k)) - m);
    f[4 + k] = __float_as_uint(__uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7540 + k)) - m);
  }
  int4 o;
  o.x = (int)__byte_perm(f[0], f[1], 0x7632);
  o.y = (int)__byte_perm(f[2], f[3], 0x7632);
  o.z = (int)__byte_perm(f[4], f[5], 0x7632);
  o.w = (int)__byte_perm(f[6], f[7], 0x7632);
  return o;
}

// The uint8 pixels travel ring -> shared memory by TMA, several tiles ahead of the converters, so that no thread ever waits on a
// global load.  The ring is described to the TMA as a 3-D tensor of 32-bit words [capacity rows][G grid rows][frame_w words]
// (the 4 pixel rows of one grid row of one frame are 4 * frame_w contiguous bytes = frame_w words): the pixels a slab needs from
// ONE image are the box [4 frames][slots grid rows][frame_w words] at (word 0, grid row q0 - b*G, ring row idx[b] + first) --
// one tensor load per image the slab touches (at most two), instead of one bulk copy per (image, frame) whose fixed cost
// (~0.2 us each, serialised per SM) made the kernel twice as slow as the bf16 version.  Grid rows past the end of the image are
// zero-filled by the TMA and never read.
//   staging layout of one tile: [image segment 0 | 1][frame f][slot = grid-row index q - qseg][4 * frame_w bytes]
constexpr int U8_MAX_STAGES = 8;
__host__ __device__ inline int u8_slots(int slab_rows, int G) { return (slab_rows + G - 2) / G + 1; }
__host__ __device__ inline int u8_box_bytes(int slab_rows, int G, int frame_w) {
  return (4 * u8_slots(slab_rows, G) * 4 * frame_w + 127) & ~127;
}
__host__ __device__ inline int u8_stage_bytes(int slab_rows, int G, int frame_w) { return 2 * u8_box_bytes(slab_rows, G, frame_w); }
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          s2u(smem_dst)),
      "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// images a slab starting at grid-matrix row R0 touches: b0 (always, if any row is inside the batch) and b1 = b0 + 1 (n1 > 0)
struct U8Plan { int b0, n0, n1, q0; };
__device__ __forceinline__ U8Plan u8_plan(const U8Src& u, int R0, int slab_rows) {
  U8Plan pl;
  const int totalq = u.rows / u.G;
  pl.q0 = R0 / u.G;
  int q1 = (R0 + slab_rows - 1) / u.G;
  if (q1 > totalq - 1) q1 = totalq - 1;
  pl.b0 = 0; pl.n0 = 0; pl.n1 = 0;
  if (q1 >= pl.q0) {
    pl.b0 = pl.q0 / u.G;
    const int b1 = q1 / u.G;
    pl.n0 = min(q1, pl.b0 * u.G + u.G - 1) - pl.q0 + 1;
    pl.n1 = b1 > pl.b0 ? q1 - b1 * u.G + 1 : 0;
  }
  return pl;
}
// ONE thread: arm `bar` with the tile's byte count and issue its one or two tensor loads; i0 / i1 = idx[b0] / idx[b0 + 1]
__device__ __forceinline__ void u8_issue(const U8Src& u, const CUtensorMap* map, const U8Plan& pl, long long i0, long long i1,
                                         uint8_t* stage, uint64_t* bar, int slots) {
  const uint32_t box = (uint32_t)(4 * slots * 4 * u.frame_w);          // bytes the TMA reports per box (zero fill included)
  const int nb = (pl.n0 > 0) + (pl.n1 > 0);
  if (nb == 0) { mb_arrive(bar); return; }
  mb_expect_tx(bar, box * nb);
  tma_load_3d(stage, map, bar, 0, pl.q0 - pl.b0 * u.G, (int)(i0 + u.first));
  if (pl.n1 > 0) tma_load_3d(stage + ((box + 127) & ~127u), map, bar, 0, 0, (int)(i1 + u.first));
}

// explicit shared-space accesses: the staging / slab pointers are carved out of the dynamic shared memory through integer
// arithmetic, so the compiler only sees GENERIC pointers -- generic loads of shared memory go through the L1 path with ~10x the
// latency of LDS (measured: the converters took 4 000 cycles per tile with them, long_scoreboard-bound)
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const int4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int J0, int NJ>
__device__ __forceinline__ void u8_store_chunks(uint32_t stage, uint32_t slab, int rl, int so, int fstride, int frame_w) {
  // chunks J0 .. J0+NJ-1 (16 bytes = 8 channels each) of slab row rl; so = staging offset of pixel (4gy, 4gx) of frame 0, or -1: zeros
  uint32_t w0[NJ], w1[NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {                       // all loads first, then the conversions
    const int j = J0 + jj;
    w0[jj] = w1[jj] = 0;
    if (so >= 0) {
      const uint32_t src = stage + (uint32_t)((j >> 1) * fstride + so + (2 * (j & 1)) * frame_w);
      w0[jj] = lds32(src);
      w1[jj] = lds32(src + frame_w);
    }
  }
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = J0 + jj;
    sts128(slab + (uint32_t)(rl * 128 + ((j ^ (rl & 7)) << 4)), cvt8_u8_bf16(w0[jj], w1[jj]));
  }
}

// NT = 128 or 256 threads (tid 0..NT-1, named barrier `bar_id`) convert one staged tile into the swizzled bf16 slab of
// `slab_rows` rows x 64 channels whose first row is grid-matrix row R0; rows >= u.rows are zero (what the TMA's out-of-bounds
// fill gave).  Thread t owns slab row t & 127 (consecutive threads -> consecutive pixels of the staging rows and the 8 distinct
// swizzle positions of a 128-byte window: conflict-free both ways) and, with 256 threads, one half of its 8 chunks; rows beyond
// 128 are shared out chunk-wise.  On return every thread's stores are fenced towards the async proxy (tcgen05.mma reads shared
// memory through it) and all NT threads have arrived.
template <int NT>
__device__ __forceinline__ void u8_convert(const U8Src& u, const uint8_t* stage_p, uint8_t* slab_p, int R0, int slab_rows,
                                           int slots, int tid, int bar_id) {
  const uint32_t stage = s2u(stage_p), slab = s2u(slab_p);
  const int rowb = 4 * u.frame_w, fstride = slots * rowb;
  const int q0 = R0 / u.G;
  const int qb1 = (q0 / u.G + 1) * u.G;                   // first grid row of the second image the slab may touch
  const int seg1 = (4 * fstride + 127) & ~127;            // its box sits behind the first image's
  const int rl0 = tid & 127;
  if (rl0 < slab_rows) {
    const int r = R0 + rl0;
    int so = -1;
    if (r < u.rows) {
      const int q = r / u.G;
      so = (q >= qb1 ? seg1 + (q - qb1) * rowb : (q - q0) * rowb) + 4 * (r - q * u.G);
    }
    if (NT == 128) {
      u8_store_chunks<0, 8>(stage, slab, rl0, so, fstride, u.frame_w);
    } else if (tid < 128) {
      u8_store_chunks<0, 4>(stage, slab, rl0, so, fstride, u.frame_w);
    } else {
      u8_store_chunks<4, 4>(stage, slab, rl0, so, fstride, u.frame_w);
    }
  }
  for (int e = tid; e < (slab_rows - 128) * 8; e += NT) {
    const int rl = 128 + (e >> 3), j = e & 7, r = R0 + rl;
    uint32_t w0 = 0, w1 = 0;
    if (r < u.rows) {
      const int q = r / u.G;
      const uint32_t src = stage + (uint32_t)((j >> 1) * fstride + (q >= qb1 ? seg1 + (q - qb1) * rowb : (q - q0) * rowb) +
                                              4 * (r - q * u.G) + (2 * (j & 1)) * u.frame_w);
      w0 = lds32(src);
      w1 = lds32(src + u.frame_w);
    }
    sts128(slab + (uint32_t)(rl * 128 + ((j ^ (rl & 7)) << 4)), cvt8_u8_bf16(w0, w1));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
}

struct SlabParams {
  GemmParams g;            // M, N, ldd, relu, out_mode, bias, D, taps_x, grid_w, shift_sign, out_map, G, V
  int taps, col_blocks;    // taps, channels / 64
  int slab_rows;           // multiple of 8, >= 128 + max_shift
  int min_shift;           // row offset of the slab start relative to m0 (0 for forward, -max_shift for dgrad)
  int stages;
  int base_offset_mode;    // 1: descriptor base_offset = (window start >> 7) & 7; 2: base_offset = 0 (address-based swizzle)
  U8Src u8;                // U8 kernels: the activation slabs are built from the uint8 frame ring (K1), tmA is unused
};

// BN = 32 (conv1, 12 tiles per SM, 16 KB of weights) compiles for two resident CTAs per SM: with many tiles the work can be
// split over 2 x 148 CTAs whose waits interleave (launch_slab uses a <= 110 KB shared-memory budget then).
constexpr int SLAB_U8_THREADS = GEMM_THREADS + 128;   // K1: four converter warps (10-13) behind the ten of the TMA version
template <int BN, bool EXT, bool U8>
__global__ void __launch_bounds__(U8 ? SLAB_U8_THREADS : GEMM_THREADS, (BN == 32 && !U8) ? 2 : 1) conv_slab_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                            const __grid_constant__ CUtensorMap tmB,
                                                                            const __grid_constant__ CUtensorMap tmA2,
                                                                            const __grid_constant__ CUtensorMap tmB2,
                                                                            const SlabParams sp) {
  const int n_cta = sp.g.dual ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool second = sp.g.dual && (int)blockIdx.x >= n_cta;
  const int cta = second ? (int)blockIdx.x - n_cta : (int)blockIdx.x;
  const CUtensorMap* mA = second ? &tmA2 : &tmA;
  const CUtensorMap* mB = second ? &tmB2 : &tmB;
  GemmParams p = sp.g;
  if (second) { p.D = sp.g.D2; p.bias = sp.g.bias2; }
  if (threadIdx.x == 0) B2RL_TRACE_AT(3, 0, 0);
  constexpr uint32_t W_TILE = BN * 128;                               // one 64-wide k-tile of the weights
  constexpr int ACC = 1;
  constexpr uint32_t TMEM_COLS = 2 * ACC * BN < 32 ? 32 : 2 * ACC * BN;
  constexpr int MAX_STAGES = 6;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k_tiles = sp.taps * sp.col_blocks;
  const uint32_t slab_block = (uint32_t)sp.slab_rows * 128;          // one 64-channel column block of a slab
  const uint32_t slab_bytes = slab_block * sp.col_blocks;
  uint8_t* sW = smem;
  uint8_t* sS = smem + (size_t)k_tiles * W_TILE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sS + (size_t)sp.stages * slab_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  // U8 (K1): uint8 staging tiles + their full / empty barriers live behind the slab ring (launch_slab_t sizes the allocation)
  const int u8_slots_ = U8 ? u8_slots(sp.slab_rows, sp.u8.G) : 0;
  const int u8_bytes = U8 ? u8_stage_bytes(sp.slab_rows, sp.u8.G, sp.u8.frame_w) : 0;
  uint8_t* sU = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));
  const int U8_STAGES = U8 ? sp.u8.stages : 1;
  uint64_t* u8_full = reinterpret_cast<uint64_t*>(sU + (size_t)U8_STAGES * u8_bytes);
  uint64_t* u8_empty = u8_full + U8_MAX_STAGES;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
  const int tiles = (p.M + GEMM_BM - 1) / GEMM_BM;

  __shared__ float s_dbias[128];                    // per-CTA bias-gradient accumulator (backward extras)
  if (threadIdx.x < 128) s_dbias[threadIdx.x] = 0.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mb_init(&tmem_full[s], 1); mb_init(&tmem_empty[s], 4); }
    mb_init(w_full, 1);
    if (U8) {
      for (int s = 0; s < U8_STAGES; ++s) { mb_init(&u8_full[s], 1); mb_init(&u8_empty[s], 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(mB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // everything above (barriers, tensor-map prefetch, TMEM allocation) overlaps the previous kernel's tail

  if (warp == 0 && elect_one()) {
   if constexpr (U8) {
    // ---------------------------------------------------------------------- K1 producer: weights, then the uint8 pixels of every
    // tile (one tensor load per image the slab touches), up to sp.u8.stages tiles ahead; idx[] is requested one tile early
    mb_expect_tx(w_full, (uint32_t)k_tiles * W_TILE);
    for (int kt = 0; kt < k_tiles; ++kt) tma_load_2d(sW + (size_t)kt * W_TILE, mB, w_full, kt * GEMM_BK, 0);
    const int nimg = sp.u8.rows / (sp.u8.G * sp.u8.G);
    U8Plan pl = u8_plan(sp.u8, cta * GEMM_BM + sp.min_shift, sp.slab_rows);
    long long i0 = 0, i1 = 0;
    if (cta < tiles && pl.n0 > 0) { i0 = __ldg(sp.u8.idx + pl.b0); i1 = __ldg(sp.u8.idx + min(pl.b0 + 1, nimg - 1)); }
    uint32_t it = 0;
    for (int tile = cta; tile < tiles; tile += n_cta, ++it) {
      const U8Plan cur = pl;
      const long long c0 = i0, c1 = i1;
      if (tile + n_cta < tiles) {
        pl = u8_plan(sp.u8, (tile + n_cta) * GEMM_BM + sp.min_shift, sp.slab_rows);
        if (pl.n0 > 0) { i0 = __ldg(sp.u8.idx + pl.b0); i1 = __ldg(sp.u8.idx + min(pl.b0 + 1, nimg - 1)); }
      }
      const int us = it % U8_STAGES;
      mb_wait(&u8_empty[us], ((it / U8_STAGES) & 1) ^ 1);
      u8_issue(sp.u8, mA, cur, c0, c1, sU + (size_t)us * u8_bytes, &u8_full[us], u8_slots_);
    }
   } else {
    // ---------------------------------------------------------------------- TMA producer
    B2RL_TRACE_AT(3, 0, 1);
    mb_expect_tx(w_full, (uint32_t)k_tiles * W_TILE);
    for (int kt = 0; kt < k_tiles; ++kt) tma_load_2d(sW + (size_t)kt * W_TILE, mB, w_full, kt * GEMM_BK, 0);
    uint32_t it = 0;
    if (!U8) {
      for (int tile = cta; tile < tiles; tile += n_cta, ++it) {
        const int s = it % sp.stages;
        B2RL_TRACE_AT(0, it, 0);
        mb_wait(&empty[s], ((it / sp.stages) & 1) ^ 1);
        B2RL_TRACE_AT(0, it, 1);
        mb_expect_tx(&full[s], slab_bytes);
        for (int cb = 0; cb < sp.col_blocks; ++cb)
          tma_load_2d(sS + (size_t)s * slab_bytes + (size_t)cb * slab_block, mA, &full[s], cb * GEMM_BK,
                      tile * GEMM_BM + sp.min_shift);
      }
    }
   }
  } else if (U8 && warp >= 10) {
    // ---------------------------------------------------------------------- K1 converters (warps 10-13): staged uint8 -> slab
    const int tid = (int)threadIdx.x - GEMM_THREADS;
    uint32_t it = 0;
    for (int tile = cta; tile < tiles; tile += n_cta, ++it) {
      const int s = it % sp.stages, us = it % U8_STAGES;
      mb_wait(&empty[s], ((it / sp.stages) & 1) ^ 1);
      mb_wait(&u8_full[us], (it / U8_STAGES) & 1);
      u8_convert<128>(sp.u8, sU + (size_t)us * u8_bytes, sS + (size_t)s * slab_bytes, tile * GEMM_BM + sp.min_shift,
                      sp.slab_rows, u8_slots_, tid, 2);
      if (tid == 0) { mb_arrive(&full[s]); mb_arrive(&u8_empty[us]); }
    }
  } else if (warp == 1 && elect_one()) {
    // ---------------------------------------------------------------------- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
    mb_wait(w_full, 0);
    B2RL_TRACE_AT(3, 0, 2);
    const uint32_t slab_lo0 = desc_lo(s2u(sS), 16), w_lo0 = desc_lo(s2u(sW), 16);
    const int taps_y = sp.taps / p.taps_x;
    uint32_t it = 0;
    for (int tile = cta; tile < tiles; tile += n_cta, ++it) {
      const uint32_t as = it & 1;
      const int s = it % sp.stages;
      B2RL_TRACE_AT(1, it, 0);
      mb_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      B2RL_TRACE_AT(1, it, 1);
      mb_wait(&full[s], (it / sp.stages) & 1);
      B2RL_TRACE_AT(1, it, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + as * (ACC * BN);
      const uint32_t slab_lo = slab_lo0 + (uint32_t)s * (slab_bytes >> 4);
      uint32_t nmma = 0, b_lo = w_lo0;                                // nmma: MMAs issued for this tile (rotates the partials)
      // taps in (dy, dx) order; the window of tap (dy, dx) starts sign*(dy*grid_w + dx) - min_shift rows into the slab
      int row_dy = -sp.min_shift;                                   // rows, for dx = 0
      for (int dy = 0; dy < taps_y; ++dy, row_dy += p.shift_sign * p.grid_w) {
        int row = row_dy;
        for (int dx = 0; dx < p.taps_x; ++dx, row += p.shift_sign) {
          uint32_t a_cb = slab_lo + (uint32_t)row * 8;              // 128 bytes per row = 8 x 16 B
          for (int cb = 0; cb < sp.col_blocks; ++cb, a_cb += slab_block >> 4) {
            uint32_t a_lo = a_cb;
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              umma_f16_lh(acc + (k % ACC) * BN, a_lo, b_lo, idesc, (nmma > 0 || k >= ACC) ? 1u : 0u);
              a_lo += 2;
              b_lo += 2;
            }
            nmma = 1;
            b_lo += (W_TILE >> 4) - 8;                              // next 64-wide k-tile of the resident weights
          }
        }
      }
      umma_commit(&empty[s]);
      umma_commit(&tmem_full[as]);
      B2RL_TRACE_AT(1, it, 3);
    }
  } else if (warp >= 2 && warp < 10) {
    const int q = warp & 3;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;                    // accumulator stage this warp group drains
    uint32_t it = 0;
    for (int tile = cta; tile < tiles; tile += n_cta, ++it)
      if ((it & 1) == grp)
        epilogue_tile<BN, ACC, EXT>(p, tile * GEMM_BM, 0, q, lane, tmem_base, it & 1, (it >> 1) & 1, true, tmem_full, tmem_empty, s_dbias, it);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (EXT && p.dbias && (int)threadIdx.x < p.dbias_mod) atomicAdd(p.dbias + threadIdx.x, s_dbias[threadIdx.x]);
  if (threadIdx.x == 0) B2RL_TRACE_AT(3, 0, 3);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Slab variant of the convolution weight gradient:  D[n, tap*C + c] += sum_r G[r, n] * X[r + shift(tap), c].
// Both operands are read as stored (MN-major, K = rows).  A CTA owns a contiguous range of 64-row k-tiles; per k-tile it
// loads the gradient rows ONCE and ONE slab of 64 + max_shift activation rows, and issues every tap's MMAs on windows of
// that slab (a K-row shift is a 128-byte step in the MN-major swizzled layout, address-based swizzle as above) into one
// TMEM accumulator per tap (taps x C fp32 columns, <= 512).  At the end each CTA stores its partial sums plainly at
// D + blockIdx.x * partial_stride (the consumer sums them: deterministic), or adds them to D with fp32 vector atomics
// when partial_stride == 0.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  int rows, n_out, C, col_blocks;
  int tap0, ntaps, taps_x, grid_w;
  int slab_rows, stages, k_tiles_per_cta, a_boxes;
  float* D;
  int ldd;
  int64_t partial_stride;   // 0: atomically accumulate into D; > 0: CTA i stores its partial sums at D + i*partial_stride
  // MMAs of one k16 step, built on the host (launch_wgrad): with one 64-channel column block per tap (C == 64) the taps of
  // one grid row (dx = 0..taps_x-1) are ONE MMA -- their windows start 1 row = 128 bytes apart in the slab, which is an
  // MN-major B operand of N = run*64 columns whose 64-element groups are LBO = 128 bytes apart (the 128-byte swizzle is a
  // function of the address, so overlapping groups read the right bytes) with adjacent TMEM accumulator columns.
  int n_runs;
  uint32_t run_off[9], run_acc[9], run_n[9];   // slab window offset (16-byte units), accumulator column, N of the MMA
  // M-stacking (n_out <= 64): the A operand's second 64-row half holds the SAME gradient columns read `stack_delta` rows away
  // (a second TMA box), so accumulator lanes 64-127 of a window at shift s hold the tap at shift s - stack_delta: with
  // stack_delta = -grid_w the windows of tap row dy also produce tap row dy + 1 -- the last tap row costs no MMAs at all
  // (conv2 / conv1: half the MMAs; conv3: 6 windows instead of 9 taps, 384 columns, ONE launch instead of two).
  // run_low[j] = first output column of the lower half of run j, or -1 (a duplicate of an upper tap: dropped).
  int stack_delta, stack_rows;
  int run_low[9];
  U8Src u8;                                    // U8 kernels: the activation slabs come from the uint8 frame ring (K1)
};

template <int TMEM_COLS, bool U8>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                             const __grid_constant__ CUtensorMap tmX,
                                                                             const WgradParams w) {
  if (threadIdx.x == 0) B2RL_TRACE_AT(3, 0, 0);
  constexpr int MAX_STAGES = 6;
  constexpr uint32_t A_BYTES = 2 * 8192;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t slab_block = (uint32_t)w.slab_rows * 128, slab_bytes = slab_block * w.col_blocks;
  const uint32_t stage_bytes = A_BYTES + slab_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)w.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  // U8 (K1): uint8 staging tiles + their full / empty barriers behind the operand ring (launch_wgrad sizes the allocation)
  const int u8_slots_ = U8 ? u8_slots(w.slab_rows, w.u8.G) : 0;
  const int u8_bytes = U8 ? u8_stage_bytes(w.slab_rows, w.u8.G, w.u8.frame_w) : 0;
  uint8_t* sU = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));
  const int U8_STAGES = U8 ? w.u8.stages : 1;
  uint64_t* u8_full = reinterpret_cast<uint64_t*>(sU + (size_t)U8_STAGES * u8_bytes);
  uint64_t* u8_empty = u8_full + U8_MAX_STAGES;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
  const int kt_total = (w.rows + GEMM_BK - 1) / GEMM_BK;
  const int kt_begin = blockIdx.x * w.k_tiles_per_cta;
  const int n_kt = max(min(kt_total, kt_begin + w.k_tiles_per_cta) - kt_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mb_init(&full[s], U8 ? 2 : 1); mb_init(&empty[s], 1); }   // U8: TMA + converters
    mb_init(tmem_full, 1);
    if (U8) {
      for (int s = 0; s < U8_STAGES; ++s) { mb_init(&u8_full[s], 1); mb_init(&u8_empty[s], 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // everything above (barriers, tensor-map prefetch, TMEM allocation) overlaps the previous kernel's tail

  if (warp == 0 && elect_one()) {
   if constexpr (U8) {
    // K1 producer: the uint8 pixels of every k-tile (one tensor load per image its slab touches, up to w.u8.stages k-tiles ahead,
    // idx[] one k-tile early) and the gradient rows
    const int nimg = w.u8.rows / (w.u8.G * w.u8.G);
    U8Plan pl = u8_plan(w.u8, kt_begin * GEMM_BK, w.slab_rows);
    long long i0 = 0, i1 = 0;
    if (n_kt > 0 && pl.n0 > 0) { i0 = __ldg(w.u8.idx + pl.b0); i1 = __ldg(w.u8.idx + min(pl.b0 + 1, nimg - 1)); }
    for (int i = 0; i < n_kt; ++i) {
      const U8Plan cur = pl;
      const long long c0 = i0, c1 = i1;
      if (i + 1 < n_kt) {
        pl = u8_plan(w.u8, (kt_begin + i + 1) * GEMM_BK, w.slab_rows);
        if (pl.n0 > 0) { i0 = __ldg(w.u8.idx + pl.b0); i1 = __ldg(w.u8.idx + min(pl.b0 + 1, nimg - 1)); }
      }
      const int us = i % U8_STAGES, s = i % w.stages;
      mb_wait(&u8_empty[us], ((i / U8_STAGES) & 1) ^ 1);
      u8_issue(w.u8, &tmX, cur, c0, c1, sU + (size_t)us * u8_bytes, &u8_full[us], u8_slots_);
      mb_wait(&empty[s], ((i / w.stages) & 1) ^ 1);
      mb_expect_tx(&full[s], (uint32_t)w.a_boxes * 8192);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      for (int g = 0; g < w.a_boxes; ++g)
        tma_load_2d(st + g * 8192, &tmG, &full[s], w.stack_delta ? 0 : g * 64, (kt_begin + i) * GEMM_BK + (g ? w.stack_delta : 0));
    }
   } else {
    B2RL_TRACE_AT(3, 0, 1);
    for (int i = 0; i < n_kt; ++i) {
      const int s = i % w.stages;
      B2RL_TRACE_AT(0, i, 0);
      mb_wait(&empty[s], ((i / w.stages) & 1) ^ 1);
      B2RL_TRACE_AT(0, i, 1);
      mb_expect_tx(&full[s], (uint32_t)w.a_boxes * 8192 + (U8 ? 0u : slab_bytes));
      uint8_t* st = smem + (size_t)s * stage_bytes;
      const int k0 = (kt_begin + i) * GEMM_BK;
      for (int g = 0; g < w.a_boxes; ++g)                                                                 // [64 k][64 n]
        tma_load_2d(st + g * 8192, &tmG, &full[s], w.stack_delta ? 0 : g * 64, k0 + (g ? w.stack_delta : 0));
      if (!U8) {
        for (int cb = 0; cb < w.col_blocks; ++cb)
          tma_load_2d(st + A_BYTES + (size_t)cb * slab_block, &tmX, &full[s], cb * GEMM_BK, k0);         // [slab_rows][64 c]
      }
    }
   }
  } else if (warp == 1 && n_kt > 0 && elect_one()) {
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(GEMM_BM >> 4) << 24);
    const bool merge = w.C == 64;
    const uint32_t a_lo0 = desc_lo(s2u(smem), 8192), b_lo0 = desc_lo(s2u(smem + A_BYTES), merge ? 128u : slab_block);
    for (int i = 0; i < n_kt; ++i) {
      const int s = i % w.stages;
      B2RL_TRACE_AT(1, i, 0);
      mb_wait(&full[s], (i / w.stages) & 1);
      B2RL_TRACE_AT(1, i, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_st = a_lo0 + (uint32_t)s * (stage_bytes >> 4), b_st = b_lo0 + (uint32_t)s * (stage_bytes >> 4);
      // runs outer (table in the constant bank, not unrolled), k16 steps inner: the issuing thread must stay lean --
      // a fully unrolled predicated table cost ~190 cycles of bookkeeping per MMA
      const uint32_t first = i > 0 ? 1u : 0u;
#pragma unroll 1
      for (int j = 0; j < w.n_runs; ++j) {
        const uint32_t acc = tmem_base + w.run_acc[j], idesc = idesc0 | ((w.run_n[j] >> 3) << 17);
        const uint32_t b_j = b_st + w.run_off[j];
        umma_f16_lh(acc, a_st, b_j, idesc, first);
#pragma unroll
        for (int k = 1; k < GEMM_BK / 16; ++k) umma_f16_lh(acc, a_st + k * 128, b_j + k * 128, idesc, 1u);
      }
      umma_commit(&empty[s]);
      B2RL_TRACE_AT(1, i, 3);
    }
    umma_commit(tmem_full);
  } else if (warp >= 2 && n_kt > 0) {
    if (U8) {
      // K1 converters (all eight epilogue warps, idle until the accumulators are complete): the activation slab of every k-tile
      // from the staged uint8 pixels
      const int tid = (int)threadIdx.x - 2 * 32;
      for (int i = 0; i < n_kt; ++i) {
        const int s = i % w.stages, us = i % U8_STAGES;
        mb_wait(&empty[s], ((i / w.stages) & 1) ^ 1);
        mb_wait(&u8_full[us], (i / U8_STAGES) & 1);
        u8_convert<256>(w.u8, sU + (size_t)us * u8_bytes, smem + (size_t)s * stage_bytes + A_BYTES, (kt_begin + i) * GEMM_BK,
                        w.slab_rows, u8_slots_, tid, 2);
        if (tid == 0) { mb_arrive(&full[s]); mb_arrive(&u8_empty[us]); }
      }
    }
    const int q = warp & 3;
    const int n = q * 32 + lane;
    if (warp == 4 && lane == 0) B2RL_TRACE_AT(2, 0, 0);
    mb_wait(tmem_full, 0);
    if (warp == 4 && lane == 0) B2RL_TRACE_AT(2, 0, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int cols = w.ntaps * w.C;
    const bool low = w.stack_delta != 0 && q >= 2;                     // lanes 64-127: the stacked copy (warp-uniform)
    const int nn = low ? n - 64 : n;
    for (int c = ((warp - 2) >> 2) * 32; c < cols; c += 64) {         // the two warp groups take alternate 32-column chunks
      int oc = w.tap0 * w.C + c;
      if (low) {
        int j = 0;
        while (j + 1 < w.n_runs && (uint32_t)c >= w.run_acc[j + 1]) ++j;
        if (w.run_low[j] < 0) continue;                                // duplicate of a tap the upper half already holds
        oc = w.run_low[j] + (c - (int)w.run_acc[j]);
      }
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      if (nn < w.n_out) {
        float* d = w.D + (int64_t)blockIdx.x * w.partial_stride + (int64_t)nn * w.ldd + oc;
        if (w.partial_stride > 0) {
          // split-K partials are stored plainly (coalesced 16-byte stores) and summed by the consumer
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        } else if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)       // 16-byte vector reductions (red.global.add.v4.f32)
            atomicAdd(reinterpret_cast<float4*>(d + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(d + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4 && lane == 0) B2RL_TRACE_AT(2, 0, 3);
  if (threadIdx.x == 0) B2RL_TRACE_AT(3, 0, 3);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [outer][inner] matrix with row stride `ld` elements, box = [box_outer][64]
static int make_map(CUtensorMap* m, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2RL_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return B2RL_ERR_CUDA; }
  return B2RL_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static bool has_ext(const GemmParams& p) { return p.mask || p.dbias || p.out_map >= 3; }

template <int BN, int STAGES, bool EXT>
static int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                         const GemmParams& p, int splits, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + (2 * STAGES + 4) * 8 + 16;
  auto k = gemm_tcgen05_kernel<BN, STAGES, EXT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN);
  int ctas = sm_count() / splits / (p.dual ? 2 : 1);                 // per operand set
  if (ctas < 1) ctas = 1;
  if (ctas > tiles) ctas = tiles;
  if (p.out_mode == 3 && ctas != tiles) {
    set_error("b2rl_gemm_splitk_bf16: tiles x splits (%d x %d) must fit one CTA per SM (%d) and splits <= 8", tiles, splits, sm_count());
    return B2RL_ERR_ARG;
  }
  dim3 grid(p.dual ? 2 * ctas : ctas, 1, splits);
  launch_pdl(k, dim3(grid), dim3(GEMM_THREADS), smem, st, ta, tb, ta2, tb2, p);
  return check_launch("b2rl_gemm_bf16");
}

// the backward extras (mask / bias gradient / scatter maps) are compiled into their own instantiation: the forward and
// weight-gradient kernels keep the lean epilogue (102 instead of 168 registers)
template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                       const GemmParams& p, int splits, cudaStream_t st) {
  return has_ext(p) ? launch_gemm_t<BN, STAGES, true>(ta, tb, ta2, tb2, p, splits, st)
                    : launch_gemm_t<BN, STAGES, false>(ta, tb, ta2, tb2, p, splits, st);
}

static int gemm_dispatch(const uint16_t* A, int a_mn, int64_t lda, int64_t a_rows, int64_t a_cols, const uint16_t* B,
                         int b_mn, int64_t ldb, int64_t b_rows, int64_t b_cols, GemmParams p, int splits, int block_n,
                         cudaStream_t st, const uint16_t* A2 = nullptr, const uint16_t* B2 = nullptr) {
  CUtensorMap ta, tb, ta2, tb2;
  int rc;
  // the tensor map describes the matrix AS STORED: [a_rows][a_cols]; box = [128|64 rows][64 cols]
  rc = make_map(&ta, A, a_cols, a_rows, lda, a_mn ? 64 : GEMM_BM);
  if (rc) return rc;
  rc = make_map(&tb, B, b_cols, b_rows, ldb, b_mn ? 64 : block_n);
  if (rc) return rc;
  const int kt_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  if (splits > kt_total) splits = kt_total;
  p.k_tiles_per_split = (kt_total + splits - 1) / splits;
  splits = (kt_total + p.k_tiles_per_split - 1) / p.k_tiles_per_split;
  ta2 = ta, tb2 = tb;
  if (p.dual) {                                                      // second operand set: same shapes and strides
    rc = make_map(&ta2, A2, a_cols, a_rows, lda, a_mn ? 64 : GEMM_BM);
    if (rc) return rc;
    rc = make_map(&tb2, B2, b_cols, b_rows, ldb, b_mn ? 64 : block_n);
    if (rc) return rc;
  }
  if (block_n == 32) return launch_gemm<32, 6>(ta, tb, ta2, tb2, p, splits, st);
  if (block_n == 64) return launch_gemm<64, 6>(ta, tb, ta2, tb2, p, splits, st);
  return launch_gemm<128, 5>(ta, tb, ta2, tb2, p, splits, st);
}

template <int BN, bool EXT, bool U8>
static int launch_slab_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                         SlabParams sp, cudaStream_t st) {
  const size_t w_bytes = (size_t)sp.taps * sp.col_blocks * BN * 128;
  const size_t slab_bytes = (size_t)sp.slab_rows * 128 * sp.col_blocks;
  const int tiles = (sp.g.M + GEMM_BM - 1) / GEMM_BM;
  static int two_cta = -1;                                           // B2RL_SLAB_2CTA=0 keeps one CTA per SM for BN = 32 too
  if (two_cta < 0) {
    const char* e = getenv("B2RL_SLAB_2CTA");
    two_cta = (e && atoi(e) == 0) ? 0 : 1;
  }
  // two CTAs per SM when the kernel was compiled for it, the tiles are plentiful and both fit (<= 110 KB each, >= 3 stages)
  const bool pair = BN == 32 && two_cta && !sp.g.dual && tiles >= 4 * sm_count() &&
                    w_bytes + 3 * slab_bytes + 2048 + (U8 ? 3 * 16 * 1024 : 0) <= 110 * 1024;
  size_t u8_extra = 0;
  size_t budget = pair ? 110 * 1024 - 2048 : 200 * 1024;
  if (U8) {
    // K1: three bf16 slabs are enough (shared -> shared conversion); everything else goes to uint8 staging tiles, each of which
    // is held for the copy latency plus the conversion
    const size_t ub = u8_stage_bytes(sp.slab_rows, sp.u8.G, sp.u8.frame_w);
    if (w_bytes + 3 * slab_bytes + 2 * ub + 512 > budget) return 1;
    int us = (int)((budget - w_bytes - 3 * slab_bytes - 512) / ub);
    if (us > U8_MAX_STAGES) us = U8_MAX_STAGES;
    sp.u8.stages = us;
    u8_extra = (size_t)us * ub + 2 * U8_MAX_STAGES * 8;
    budget = w_bytes + 3 * slab_bytes;
  }
  if (w_bytes + 2 * slab_bytes > budget) return 1;                   // does not fit: caller falls back to tap addressing
  int stages = (int)((budget - w_bytes) / slab_bytes);
  if (stages > 6) stages = 6;
  sp.stages = stages;
  const size_t smem = 1024 + w_bytes + stages * slab_bytes + (2 * 6 + 5) * 8 + 16 + 144 + u8_extra;
  auto k = conv_slab_tcgen05_kernel<BN, EXT, U8>;
  static size_t attr = 0;
  if (attr < smem) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  int ctas = sp.g.dual ? sm_count() / 2 : (pair ? 2 * sm_count() : sm_count());   // per operand set
  if (ctas > tiles) ctas = tiles;
  launch_pdl(k, dim3(sp.g.dual ? 2 * ctas : ctas), dim3(U8 ? SLAB_U8_THREADS : GEMM_THREADS), smem, st, ta, tb, ta2, tb2, sp);
  return check_launch("b2rl_conv_gemm_bf16(slab)");
}

template <int BN>
static int launch_slab(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                       SlabParams sp, cudaStream_t st) {
  if (sp.u8.frames) {                                                // K1: conv1 forward straight from the uint8 ring
    if constexpr (BN == 32) {
      if (!has_ext(sp.g) && !sp.g.dual && sp.col_blocks == 1) return launch_slab_t<32, false, true>(ta, tb, ta2, tb2, sp, st);
    }
    set_error("the uint8-ring producer serves block_n 32, 64 channels, no backward extras, no dual launch");
    return B2RL_ERR_ARG;
  }
  return has_ext(sp.g) ? launch_slab_t<BN, true, false>(ta, tb, ta2, tb2, sp, st)
                       : launch_slab_t<BN, false, false>(ta, tb, ta2, tb2, sp, st);
}

template <int TMEM_COLS, bool U8 = false>
static int launch_wgrad(const CUtensorMap& tg, const CUtensorMap& tx, WgradParams w, cudaStream_t st, int* n_ctas = nullptr) {
  const size_t slab_bytes = (size_t)w.slab_rows * 128 * w.col_blocks, stage = 16384 + slab_bytes;
  size_t u8_extra = 0;
  int stages = (int)((200 * 1024) / stage);
  if (U8) {                                                          // K1: four operand stages, the rest for uint8 staging tiles
    const size_t ub = u8_stage_bytes(w.slab_rows, w.u8.G, w.u8.frame_w);
    if (4 * stage + 2 * ub + 512 > 200 * 1024) return 1;
    stages = 4;
    int us = (int)((200 * 1024 - 4 * stage - 512) / ub);
    if (us > U8_MAX_STAGES) us = U8_MAX_STAGES;
    w.u8.stages = us;
    u8_extra = (size_t)us * ub + 2 * U8_MAX_STAGES * 8 + 144;
  }
  if (stages > 6) stages = 6;
  if (stages < 2) return 1;
  w.stages = stages;
  const size_t smem = 1024 + stages * stage + (2 * 6 + 2) * 8 + 16 + u8_extra;
  auto k = conv_wgrad_tcgen05_kernel<TMEM_COLS, U8>;
  static size_t attr = 0;
  if (attr < smem) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  {
    const bool merge = w.C == 64;
    int tap = w.tap0, n = 0;
    uint32_t acc = 0;
    const int tap_end = w.tap0 + w.ntaps;
    while (tap < tap_end && n < 9) {
      const int dy = tap / w.taps_x, dx = tap - dy * w.taps_x;
      int run = 1;
      if (merge) {
        run = w.taps_x - dx;
        if (run > tap_end - tap) run = tap_end - tap;
      }
      w.run_off[n] = (uint32_t)(dy * w.grid_w + dx) * 8;
      w.run_acc[n] = acc;
      w.run_n[n] = (uint32_t)(run * w.C);
      w.run_low[n] = (w.stack_delta && dy == w.stack_rows - 2) ? ((dy + 1) * w.taps_x + dx) * w.C : -1;
      acc += run * w.C;
      tap += run;
      ++n;
    }
    if (tap < tap_end) return 1;
    w.n_runs = n;
  }
  const int kt_total = (w.rows + GEMM_BK - 1) / GEMM_BK;
  // split-K over all SMs (the MMAs are shared-memory-read bound per SM, so the work must be spread); every CTA ends with
  // n_out x cols / 4 vector reductions (red.global.add.v4.f32) into the same small D
  int ctas = kt_total / 4;
  if (ctas > sm_count()) ctas = sm_count();
  static int cap = -1;                                               // tunable: B2RL_WGRAD_CTAS caps the split-K width
  if (cap < 0) {
    const char* e = getenv("B2RL_WGRAD_CTAS");
    cap = e ? atoi(e) : 0;
  }
  if (cap > 0 && ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  w.k_tiles_per_cta = (kt_total + ctas - 1) / ctas;
  ctas = (kt_total + w.k_tiles_per_cta - 1) / w.k_tiles_per_cta;
  if (n_ctas) *n_ctas = ctas;
  launch_pdl(k, dim3(ctas), dim3(GEMM_THREADS), smem, st, tg, tx, w);
  return check_launch("b2rl_conv_gemm_bf16(wgrad slab)");
}

}  // namespace b2rl

#ifdef B2RL_TRACE
extern "C" int b2rl_debug_set_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(b2rl::g_trace_buf, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#endif

using namespace b2rl;

static int g_wgrad_partials = 0;          // set by b2rl_conv_wgrad_partials for the duration of one call
static float* g_partial_buf = nullptr;
static int64_t g_partial_stride = 0;
static int g_partial_count = 0;
static int g_use_slab = 2;   // 2: shifted windows with base_offset 0 -- the 128B swizzle is a pure function of the smem address (verified on B200)
extern "C" void b2rl_set_conv_slab(int32_t on) { g_use_slab = on; }

static int check_common(const void* A, const void* B, const void* D, int64_t lda, int64_t ldb, int M, int N, int K,
                        int out_mode, int splits, int block_n, int b_mn, int relu) {
  B2RL_REQUIRE(A && B && D, "null pointer");
  B2RL_REQUIRE(M > 0 && N > 0 && K > 0, "bad shape");
  B2RL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "operand row strides must be multiples of 8 elements (16 bytes)");
  B2RL_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) % 16 == 0, "operands must be 16-byte aligned");
  B2RL_REQUIRE(out_mode >= 0 && out_mode <= 2, "out_mode 0 (bf16) | 1 (fp32) | 2 (fp32 atomic add)");
  B2RL_REQUIRE(block_n == 32 || block_n == 64 || block_n == 128, "block_n must be 32, 64 or 128");
  B2RL_REQUIRE(!(b_mn && block_n < 64), "MN-major B needs block_n >= 64");
  B2RL_REQUIRE(splits >= 1 && (splits == 1 || out_mode == 2), "split-K needs out_mode 2 (atomic fp32 accumulation)");
  B2RL_REQUIRE(!(relu && out_mode == 2), "ReLU cannot be fused into an atomic accumulation");
  return B2RL_OK;
}

extern "C" int b2rl_gemm_bf16(const uint16_t* A, int32_t a_mn, int64_t lda, const uint16_t* B, int32_t b_mn, int64_t ldb,
                              void* D, int64_t ldd, int32_t M, int32_t N, int32_t K, const float* bias, int32_t relu,
                              int32_t out_mode, int32_t splits, int32_t block_n, void* stream) {
  int rc = check_common(A, B, D, lda, ldb, M, N, K, out_mode, splits, block_n, b_mn, relu);
  if (rc) return rc;
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.ldd = (int)ldd;
  p.a_mn = a_mn; p.b_mn = b_mn; p.relu = relu; p.out_mode = out_mode; p.bias = bias; p.D = D;
  p.taps_x = 1; p.shift_sign = 1;
  // K-major: stored [rows][K]; MN-major: stored [K][rows]
  return gemm_dispatch(A, a_mn, lda, a_mn ? K : M, a_mn ? M : K, B, b_mn, ldb, b_mn ? K : N, b_mn ? N : K, p, splits,
                       block_n, (cudaStream_t)stream);
}

// Convolution over a G x G grid as a shifted-row GEMM (see the header of this file).
//   mode 0 (forward / dgrad):  D[r, :] = sum_taps  X[r + shift(tap), :] * W[:, tap*C .. tap*C+C]^T
//        X: [rows][C] bf16 (C multiple of 64), W: [N][taps*C] bf16 K-major, shift(tap) = sign*((tap/taps_x)*grid_w + tap%taps_x)
//   mode 1 (wgrad):            D[n, tap*C + c] (+)= sum_r G[r, n] * X[r + shift(tap), c]
//        G: [rows][N_out] bf16, X: [rows][C] bf16 (C multiple of block_n), D: fp32 [N_out][taps*C], atomic accumulation
static int apply_ext(GemmParams& p, const b2rl_bwd_epilogue* ext, int N) {
  B2RL_REQUIRE(!ext->mask || (ext->mask_ld % 8 == 0 && reinterpret_cast<uintptr_t>(ext->mask) % 16 == 0 && N % 32 == 0),
               "mask rows must be 16-byte aligned and N a multiple of 32");
  B2RL_REQUIRE(!ext->dbias || (ext->dbias_mod >= 0 && ext->dbias_mod <= 128), "dbias_mod must be in [0, 128] (0: one bias per column)");
  B2RL_REQUIRE(p.out_map < 3 || (ext->sub_c > 0 && ext->sub_c % 32 == 0 && N % ext->sub_c == 0),
               "scatter maps need sub_c a multiple of 32 that divides N");
  p.mask = reinterpret_cast<const __nv_bfloat16*>(ext->mask);
  p.mask_ld = ext->mask_ld;
  p.dbias = ext->dbias;
  p.dbias_mod = ext->dbias ? ext->dbias_mod : 1;
  p.sub_c = ext->sub_c > 0 ? ext->sub_c : 32;
  return B2RL_OK;
}

static int conv_gemm_impl(int32_t mode, const uint16_t* X, int64_t rows, int32_t C, const uint16_t* W_or_G, int32_t n_out,
                          int32_t taps, int32_t taps_x, int32_t grid_w, int32_t shift_sign, void* D, int64_t ldd,
                          const float* bias, int32_t relu, int32_t out_mode, int32_t out_map, int32_t G, int32_t V,
                          int32_t splits, int32_t block_n, void* stream, const uint16_t* X2 = nullptr,
                          const uint16_t* W2 = nullptr, void* D2 = nullptr, const float* bias2 = nullptr,
                          const b2rl_bwd_epilogue* ext = nullptr) {
  B2RL_REQUIRE(mode == 0 || mode == 1, "mode 0 (forward/dgrad) or 1 (wgrad)");
  B2RL_REQUIRE(rows > 0 && C > 0 && taps > 0 && taps_x > 0 && n_out > 0, "bad shape");
  B2RL_REQUIRE(out_map >= 0 && out_map <= (ext ? 4 : 2), "bad out_map");
  GemmParams p = {};
  p.relu = relu; p.out_mode = out_mode; p.bias = bias; p.D = D; p.ldd = (int)ldd;
  p.taps_x = taps_x; p.grid_w = grid_w; p.shift_sign = shift_sign;
  p.out_map = out_map; p.G = G; p.V = V;
  p.dual = X2 != nullptr; p.D2 = D2; p.bias2 = bias2;
  if (ext) {
    int rc0 = apply_ext(p, ext, n_out);
    if (rc0) return rc0;
  }
  if (mode == 0) {
    B2RL_REQUIRE(C % 64 == 0, "forward/dgrad needs channels in multiples of 64");
    const int K = taps * C;
    int rc = check_common(X, W_or_G, D, C, K, (int)rows, n_out, K, out_mode, splits, block_n, 0, relu);
    if (rc) return rc;
    p.M = (int)rows; p.N = n_out; p.K = K; p.a_mn = 0; p.b_mn = 0; p.a_tap_tiles = C / 64;
    if (g_use_slab && splits == 1 && n_out <= block_n && out_mode != 2) {
      const int max_shift = ((taps - 1) / taps_x) * grid_w + (taps - 1) % taps_x;
      SlabParams sp = {};
      sp.g = p; sp.taps = taps; sp.col_blocks = C / 64;
      sp.slab_rows = (GEMM_BM + max_shift + 7) / 8 * 8;
      sp.min_shift = shift_sign > 0 ? 0 : -max_shift;
      sp.base_offset_mode = g_use_slab;
      if (sp.slab_rows <= 256) {
        CUtensorMap ta, tb, ta2, tb2;
        rc = make_map(&ta, X, C, rows, C, sp.slab_rows);            // box [slab_rows][64]
        if (rc) return rc;
        rc = make_map(&tb, W_or_G, K, n_out, K, block_n);           // box [block_n][64]
        if (rc) return rc;
        ta2 = ta, tb2 = tb;
        if (p.dual) {
          rc = make_map(&ta2, X2, C, rows, C, sp.slab_rows);
          if (rc) return rc;
          rc = make_map(&tb2, W2, K, n_out, K, block_n);
          if (rc) return rc;
        }
        int r2 = block_n == 32 ? launch_slab<32>(ta, tb, ta2, tb2, sp, (cudaStream_t)stream)
                 : block_n == 64 ? launch_slab<64>(ta, tb, ta2, tb2, sp, (cudaStream_t)stream)
                                 : launch_slab<128>(ta, tb, ta2, tb2, sp, (cudaStream_t)stream);
        if (r2 <= 0) return r2;                                      // launched (0) or failed (<0); 1 = does not fit
      }
    }
    return gemm_dispatch(X, 0, C, rows, C, W_or_G, 0, K, n_out, K, p, splits, block_n, (cudaStream_t)stream, X2, W2);
  }
  B2RL_REQUIRE(!p.dual, "the dual launch is for forward / dgrad GEMMs");
  B2RL_REQUIRE(C % block_n == 0, "wgrad needs channels in multiples of block_n");
  B2RL_REQUIRE(n_out % 8 == 0, "wgrad needs n_out in multiples of 8");
  int rc = check_common(W_or_G, X, D, n_out, C, n_out, taps * C, (int)rows, out_mode, splits, block_n, 1, relu);
  if (rc) return rc;
  B2RL_REQUIRE(out_mode == 2, "wgrad accumulates with out_mode 2");
  if (g_use_slab && C % 64 == 0 && C <= 128 && n_out <= 128 && shift_sign > 0) {
    int max_shift = ((taps - 1) / taps_x) * grid_w + (taps - 1) % taps_x;
    WgradParams w = {};
    w.rows = (int)rows; w.n_out = n_out; w.C = C; w.col_blocks = C / 64; w.taps_x = taps_x; w.grid_w = grid_w;
    w.a_boxes = n_out > 64 ? 2 : 1;
    // M-stacking (see WgradParams): with at most 64 output channels the second half of the 128 accumulator lanes computes the
    // last tap row from the windows of the row before it (B2RL_WGRAD_STACK=0: every tap its own window, as in round 1)
    static int stack_on = -1;
    if (stack_on < 0) {
      const char* e = getenv("B2RL_WGRAD_STACK");
      stack_on = (e && atoi(e) == 0) ? 0 : 1;
    }
    const int taps_y = taps / taps_x;
    int win_taps = taps;
    if (stack_on && n_out <= 64 && taps_y >= 2 && taps_y * taps_x == taps) {
      w.stack_delta = -grid_w; w.stack_rows = taps_y; w.a_boxes = 2;
      win_taps = (taps_y - 1) * taps_x;
      max_shift = (taps_y - 2) * grid_w + taps_x - 1;
    }
    w.slab_rows = (GEMM_BK + max_shift + 7) / 8 * 8;
    w.D = reinterpret_cast<float*>(D); w.ldd = (int)ldd;
    if (g_wgrad_partials) { w.D = g_partial_buf; w.partial_stride = g_partial_stride; }
    CUtensorMap tg, tx;
    rc = make_map(&tg, W_or_G, n_out, rows, n_out, 64);             // gradient rows: box [64 k][64 n]
    if (rc) return rc;
    rc = make_map(&tx, X, C, rows, C, w.slab_rows);                 // activation slab: box [slab_rows][64 c]
    if (rc) return rc;
    const int per_launch = 512 / C;                                 // taps whose accumulators fit in TMEM together
    if (w.stack_delta && win_taps > per_launch) {                   // (does not happen for the NatureConvBody shapes)
      w.stack_delta = 0; w.stack_rows = 0; w.a_boxes = 1; win_taps = taps;
      w.slab_rows = (GEMM_BK + ((taps - 1) / taps_x) * grid_w + (taps - 1) % taps_x + 7) / 8 * 8;
      rc = make_map(&tx, X, C, rows, C, w.slab_rows);
      if (rc) return rc;
    }
    int r2 = 0;
    for (int t0 = 0; t0 < win_taps && r2 == 0; t0 += per_launch) {
      w.tap0 = t0;
      w.ntaps = win_taps - t0 < per_launch ? win_taps - t0 : per_launch;
      const int cols = w.ntaps * C;
      r2 = cols <= 64 ? launch_wgrad<64>(tg, tx, w, (cudaStream_t)stream, &g_partial_count)
           : cols <= 128 ? launch_wgrad<128>(tg, tx, w, (cudaStream_t)stream, &g_partial_count)
           : cols <= 256 ? launch_wgrad<256>(tg, tx, w, (cudaStream_t)stream, &g_partial_count)
                         : launch_wgrad<512>(tg, tx, w, (cudaStream_t)stream, &g_partial_count);
    }
    if (r2 <= 0) return r2;
  }
  p.M = n_out; p.N = taps * C; p.K = (int)rows; p.a_mn = 1; p.b_mn = 1; p.b_tap_tiles = C / block_n;
  return gemm_dispatch(W_or_G, 1, n_out, rows, n_out, X, 1, C, rows, C, p, splits, block_n, (cudaStream_t)stream);
}

extern "C" int b2rl_conv_gemm_bf16(int32_t mode, const uint16_t* X, int64_t rows, int32_t C, const uint16_t* W_or_G,
                                   int32_t n_out, int32_t taps, int32_t taps_x, int32_t grid_w, int32_t shift_sign,
                                   void* D, int64_t ldd, const float* bias, int32_t relu, int32_t out_mode,
                                   int32_t out_map, int32_t G, int32_t V, int32_t splits, int32_t block_n, void* stream) {
  return conv_gemm_impl(mode, X, rows, C, W_or_G, n_out, taps, taps_x, grid_w, shift_sign, D, ldd, bias, relu, out_mode,
                        out_map, G, V, splits, block_n, stream);
}

// The same forward / dgrad convolution for TWO independent operand sets (online and target network) in one launch:
// D = conv(X, W) + bias and D2 = conv(X2, W2) + bias2, identical shapes.  Half of the CTAs work on each set.
extern "C" int b2rl_conv_gemm_dual_bf16(const uint16_t* X, const uint16_t* X2, int64_t rows, int32_t C, const uint16_t* W,
                                        const uint16_t* W2, int32_t n_out, int32_t taps, int32_t taps_x, int32_t grid_w,
                                        int32_t shift_sign, void* D, void* D2, int64_t ldd, const float* bias,
                                        const float* bias2, int32_t relu, int32_t out_mode, int32_t out_map, int32_t G,
                                        int32_t V, int32_t block_n, void* stream) {
  B2RL_REQUIRE(X2 && W2 && D2, "null pointer in the second operand set");
  B2RL_REQUIRE((reinterpret_cast<uintptr_t>(X2) | reinterpret_cast<uintptr_t>(W2)) % 16 == 0, "operands must be 16-byte aligned");
  B2RL_REQUIRE((bias == nullptr) == (bias2 == nullptr), "both or neither bias");
  return conv_gemm_impl(0, X, rows, C, W, n_out, taps, taps_x, grid_w, shift_sign, D, ldd, bias, relu, out_mode, out_map, G,
                        V, 1, block_n, stream, X2, W2, D2, bias2);
}

// D = A B^T + bias and D2 = A2 B2^T + bias2 (K-major operands, identical shapes) in one launch
extern "C" int b2rl_gemm_dual_bf16(const uint16_t* A, const uint16_t* A2, int64_t lda, const uint16_t* B, const uint16_t* B2,
                                   int64_t ldb, void* D, void* D2, int64_t ldd, int32_t M, int32_t N, int32_t K,
                                   const float* bias, const float* bias2, int32_t relu, int32_t out_mode, int32_t block_n,
                                   void* stream) {
  int rc = check_common(A, B, D, lda, ldb, M, N, K, out_mode, 1, block_n, 0, relu);
  if (rc) return rc;
  B2RL_REQUIRE(A2 && B2 && D2, "null pointer in the second operand set");
  B2RL_REQUIRE((reinterpret_cast<uintptr_t>(A2) | reinterpret_cast<uintptr_t>(B2)) % 16 == 0, "operands must be 16-byte aligned");
  B2RL_REQUIRE(out_mode != 2, "the dual launch stores (bf16 or fp32), it does not accumulate");
  B2RL_REQUIRE((bias == nullptr) == (bias2 == nullptr), "both or neither bias");
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.ldd = (int)ldd;
  p.relu = relu; p.out_mode = out_mode; p.bias = bias; p.D = D;
  p.taps_x = 1; p.shift_sign = 1;
  p.dual = 1; p.D2 = D2; p.bias2 = bias2;
  return gemm_dispatch(A, 0, lda, M, K, B, 0, ldb, N, K, p, 1, block_n, (cudaStream_t)stream, A2, B2);
}

// dgrad convolution with the backward extras fused into the epilogue (ReLU mask, bias gradient, grid scatter): D = mask .*
// conv_transpose(G_rows, W), see b2rl_bwd_epilogue in the header.  bf16 output.
extern "C" int b2rl_conv_gemm_bwd_bf16(const uint16_t* G_rows, int64_t rows, int32_t C, const uint16_t* W, int32_t n_out,
                                       int32_t taps, int32_t taps_x, int32_t grid_w, void* D, int64_t ldd, int32_t out_map,
                                       int32_t G, int32_t V, const b2rl_bwd_epilogue* ext, int32_t block_n, void* stream) {
  B2RL_REQUIRE(ext, "null epilogue description");
  B2RL_REQUIRE(out_map == 0 || out_map == 3, "dgrad output map 0 (same grid) or 3 (space-to-depth(2) -> grid)");
  return conv_gemm_impl(0, G_rows, rows, C, W, n_out, taps, taps_x, grid_w, -1, D, ldd, nullptr, 0, 0, out_map, G, V, 1,
                        block_n, stream, nullptr, nullptr, nullptr, nullptr, ext);
}

// D = mask .* (A B^T) with A [M][K] K-major and B K-major ([N][K]) or MN-major ([K][N]); same epilogue extras (fc4 dgrad)
extern "C" int b2rl_gemm_bwd_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int32_t b_mn, int64_t ldb, void* D,
                                  int64_t ldd, int32_t M, int32_t N, int32_t K, int32_t out_map, int32_t G, int32_t V,
                                  const b2rl_bwd_epilogue* ext, int32_t block_n, void* stream) {
  B2RL_REQUIRE(ext, "null epilogue description");
  B2RL_REQUIRE(out_map == 0 || out_map == 4, "output map 0 (plain) or 4 (per-image positions -> grid)");
  int rc = check_common(A, B, D, lda, ldb, M, N, K, 0, 1, block_n, b_mn, 0);
  if (rc) return rc;
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.ldd = (int)ldd;
  p.a_mn = 0; p.b_mn = b_mn; p.out_mode = 0; p.D = D;
  p.taps_x = 1; p.shift_sign = 1;
  p.out_map = out_map; p.G = G; p.V = V;
  rc = apply_ext(p, ext, N);
  if (rc) return rc;
  return gemm_dispatch(A, 0, lda, M, K, B, b_mn, ldb, b_mn ? K : N, b_mn ? N : K, p, 1, block_n, (cudaStream_t)stream);
}

// Weight gradient as split-K PARTIALS: partial i (one per CTA, n_partials_host of them, at most 148) is stored at
// partials + i * n_out * taps * C; the consumer (b2rl_nature_unpack_grads) sums them.  Replaces ~1M fp32 atomics per
// layer by coalesced stores.
extern "C" int b2rl_conv_wgrad_partials(const uint16_t* X, int64_t rows, int32_t C, const uint16_t* G, int32_t n_out,
                                        int32_t taps, int32_t taps_x, int32_t grid_w, float* partials,
                                        int32_t* n_partials_host, void* stream) {
  B2RL_REQUIRE(partials && n_partials_host, "null pointer");
  B2RL_REQUIRE(g_use_slab && C % 64 == 0 && C <= 128 && n_out <= 128, "partials need the slab wgrad kernel");
  g_wgrad_partials = 1; g_partial_buf = partials; g_partial_stride = (int64_t)n_out * taps * C; g_partial_count = 0;
  int rc = b2rl_conv_gemm_bf16(1, X, rows, C, G, n_out, taps, taps_x, grid_w, 1, partials, (int64_t)taps * C, nullptr, 0, 2,
                               0, 0, 0, 1, C == 128 ? 128 : 64, stream);
  g_wgrad_partials = 0;
  *n_partials_host = g_partial_count;
  return rc;
}

// D = act(A B^T + bias) in bf16 with split-K over `splits` CTAs per output tile and an in-kernel fix-up (out_mode 3): ONE
// launch instead of zero-fill + atomic split-K + bias/activation pass.  A [M][K], B [N][K] K-major bf16.  ws: fp32
// [splits][ceil(M/128)*128][ceil(N/block_n)*block_n] scratch; counters: int32 [tiles], zero-initialised once (the kernel
// re-arms them).  Launches on different streams must not share ws / counters.  fc4 of NatureConvBody (network_bodies.py:33).
extern "C" int b2rl_gemm_splitk_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, void* D, int64_t ldd,
                                     int32_t M, int32_t N, int32_t K, const float* bias, int32_t relu, int32_t splits,
                                     int32_t block_n, float* ws, int32_t* counters, void* stream) {
  B2RL_REQUIRE(A && B && D && ws && counters, "null pointer");
  B2RL_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1 && splits <= 8, "bad shape (1 <= splits <= 8)");
  B2RL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "operand row strides must be multiples of 8 elements (16 bytes)");
  B2RL_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(ws)) % 16 == 0,
               "operands and scratch must be 16-byte aligned");
  B2RL_REQUIRE(block_n == 32 || block_n == 64 || block_n == 128, "block_n must be 32, 64 or 128");
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.ldd = (int)ldd;
  p.relu = relu; p.out_mode = 3; p.bias = bias; p.D = D;
  p.taps_x = 1; p.shift_sign = 1;
  p.ws = ws; p.counters = counters;
  p.ws_rows = (M + GEMM_BM - 1) / GEMM_BM * GEMM_BM;
  p.ws_ld = (N + block_n - 1) / block_n * block_n;
  return gemm_dispatch(A, 0, lda, M, K, B, 0, ldb, N, K, p, splits, block_n, (cudaStream_t)stream);
}


// ---------------------------------------------------------------------------------------------------------------
// K1 entry points: conv1 of NatureConvBody (network_bodies.py:27; 8x8 / stride 4 over `history` stacked frames = 2x2 taps
// over the space-to-depth(4) grid) reading the sampled frame stacks STRAIGHT FROM THE uint8 REPLAY RING (replay.py:124-134)
// -- no materialised batch.  frames: ring [capacity][row_bytes]; idx: int64 [batch] sampled ring indices; `first`: ring row
// of the oldest stacked frame relative to idx[b] (-(history-1) for the state, n_step-(history-1) for the next state).
// The frame values enter as exact integers 0..255; ImageNormalizer's 1/255 is folded into W (b2rl_nature_pack_weights).
// ---------------------------------------------------------------------------------------------------------------
// 3-D tensor map over the uint8 ring as 32-bit words: [capacity][G grid rows][frame_w words], box [4 frames][slots][frame_w]
static int make_ring_map(CUtensorMap* m, const uint8_t* frames, int64_t capacity, int64_t row_bytes, int frame_w, int slots) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2RL_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)frame_w, (cuuint64_t)(frame_w / 4), (cuuint64_t)capacity};
  cuuint64_t gstr[2] = {(cuuint64_t)4 * frame_w, (cuuint64_t)row_bytes};
  cuuint32_t box[3] = {(cuuint32_t)frame_w, (cuuint32_t)slots, 4u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(frames), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (uint8 ring) failed (%d)", (int)r); return B2RL_ERR_CUDA; }
  return B2RL_OK;
}

static int u8_src(U8Src& u, const uint8_t* frames, const int64_t* idx, int32_t first, int64_t row_bytes, int32_t frame_w,
                  int32_t batch, int32_t history) {
  B2RL_REQUIRE(frames && idx, "null pointer");
  B2RL_REQUIRE(history == 4, "the uint8-ring producer packs 4 frames x 16 pixels into the 64 channels of one slab column block");
  B2RL_REQUIRE(frame_w > 0 && frame_w % 4 == 0 && row_bytes % frame_w == 0 && row_bytes / frame_w == frame_w,
               "square frames with a side that is a multiple of 4");
  B2RL_REQUIRE(reinterpret_cast<uintptr_t>(frames) % 16 == 0 && row_bytes % 16 == 0 && (4 * frame_w) % 16 == 0 && frame_w <= 256,
               "ring rows and grid rows must be 16-byte aligned (TMA strides), frames at most 256 pixels wide");
  B2RL_REQUIRE(batch > 0 && (int64_t)batch * (frame_w / 4) * (frame_w / 4) < (1LL << 31), "bad batch");
  u.frames = frames; u.idx = idx; u.row_bytes = row_bytes; u.first = first; u.frame_w = frame_w; u.G = frame_w / 4;
  u.rows = batch * u.G * u.G;
  return B2RL_OK;
}

// D = act(conv1(frames) + bias): rows (b, gy, gx) of the G x G grid, n_out <= 32 output channels, bf16, through the output row
// maps of b2rl_conv_gemm_bf16 (out_map 1 = space-to-depth(2) rows for conv2, valid V x V).  W: [n_out][4 taps * 64] bf16.
extern "C" int b2rl_conv1_u8_fwd(const uint8_t* frames, int64_t capacity, const int64_t* idx, int32_t first, int64_t row_bytes,
                                 int32_t frame_w, int32_t batch, int32_t history, const uint16_t* W, int32_t n_out, void* D, int64_t ldd,
                                 const float* bias, int32_t relu, int32_t out_map, int32_t V, void* stream) {
  SlabParams sp = {};
  int rc = u8_src(sp.u8, frames, idx, first, row_bytes, frame_w, batch, history);
  if (rc) return rc;
  B2RL_REQUIRE(W && D, "null pointer");
  B2RL_REQUIRE(n_out > 0 && n_out <= 32 && out_map >= 0 && out_map <= 2, "n_out <= 32, out_map 0..2");
  B2RL_REQUIRE(reinterpret_cast<uintptr_t>(W) % 16 == 0, "operands must be 16-byte aligned");
  const int G = sp.u8.G, C = 64, taps = 4, K = taps * C;
  GemmParams p = {};
  p.M = sp.u8.rows; p.N = n_out; p.K = K; p.ldd = (int)ldd;
  p.relu = relu; p.out_mode = 0; p.bias = bias; p.D = D;
  p.taps_x = 2; p.grid_w = G; p.shift_sign = 1; p.a_tap_tiles = 1;
  p.out_map = out_map; p.G = G; p.V = V;
  const int max_shift = G + 1;
  sp.g = p; sp.taps = taps; sp.col_blocks = 1;
  sp.slab_rows = (GEMM_BM + max_shift + 7) / 8 * 8;
  sp.min_shift = 0;
  sp.base_offset_mode = 2;
  B2RL_REQUIRE(sp.slab_rows <= 256, "frame too wide for one slab");
  B2RL_REQUIRE(capacity > 0, "bad capacity");
  CUtensorMap tb, tr;
  rc = make_map(&tb, W, K, n_out, K, 32);
  if (rc) return rc;
  rc = make_ring_map(&tr, frames, capacity, row_bytes, frame_w, u8_slots(sp.slab_rows, G));
  if (rc) return rc;
  int r2 = launch_slab<32>(tr, tb, tr, tb, sp, (cudaStream_t)stream);
  if (r2 > 0) { set_error("b2rl_conv1_u8_fwd: the slab does not fit in shared memory"); return B2RL_ERR_ARG; }
  return r2;
}

// split-K partials of conv1's weight gradient dW[n][tap*64 + c] = sum_r G[r][n] * x[r + shift(tap)][c] with x read from the
// ring as above (G_rows: bf16 [batch*G*G][n_out], the masked output gradient on conv1's grid).  Same output contract as
// b2rl_conv_wgrad_partials.
extern "C" int b2rl_conv1_u8_wgrad_partials(const uint8_t* frames, int64_t capacity, const int64_t* idx, int32_t first,
                                            int64_t row_bytes, int32_t frame_w, int32_t batch, int32_t history, const uint16_t* G_rows,
                                            int32_t n_out, float* partials, int32_t* n_partials_host, void* stream) {
  WgradParams w = {};
  int rc = u8_src(w.u8, frames, idx, first, row_bytes, frame_w, batch, history);
  if (rc) return rc;
  B2RL_REQUIRE(G_rows && partials && n_partials_host, "null pointer");
  B2RL_REQUIRE(n_out > 0 && n_out <= 64 && n_out % 8 == 0, "n_out <= 64, multiple of 8");
  B2RL_REQUIRE(reinterpret_cast<uintptr_t>(G_rows) % 16 == 0, "operands must be 16-byte aligned");
  const int G = w.u8.G, C = 64, taps = 4;
  w.rows = w.u8.rows; w.n_out = n_out; w.C = C; w.col_blocks = 1; w.taps_x = 2; w.grid_w = G;
  // M-stacking (WgradParams): windows of tap row 0 only; accumulator lanes 64-127 (gradient rows read G rows earlier) give row 1
  w.stack_delta = -G; w.stack_rows = 2; w.a_boxes = 2;
  w.slab_rows = (GEMM_BK + 1 + 7) / 8 * 8;
  w.D = partials; w.ldd = taps * C; w.partial_stride = (int64_t)n_out * taps * C;
  w.tap0 = 0; w.ntaps = 2;
  B2RL_REQUIRE(w.slab_rows <= 256, "frame too wide for one slab");
  B2RL_REQUIRE(capacity > 0, "bad capacity");
  CUtensorMap tg, tr;
  rc = make_map(&tg, G_rows, n_out, w.rows, n_out, 64);
  if (rc) return rc;
  rc = make_ring_map(&tr, frames, capacity, row_bytes, frame_w, u8_slots(w.slab_rows, G));
  if (rc) return rc;
  int n = 0;
  int r2 = launch_wgrad<128, true>(tg, tr, w, (cudaStream_t)stream, &n);
  if (r2 > 0) { set_error("b2rl_conv1_u8_wgrad_partials: the slab does not fit in shared memory"); return B2RL_ERR_ARG; }
  *n_partials_host = n;
  return r2;
}
