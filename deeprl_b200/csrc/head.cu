// head.cu -- value heads with a handful of outputs (VanillaNet / DuelingNet, network_heads.py:11-37) on CUDA cores:
// the contraction is [B x 512] x [512 x A] with A = number of actions (4..18): far too narrow for a tensor-core tile,
// and in eager form it costs ~12 launches per update (weight / bias casts, GEMM, float cast, dueling combine, and their
// backward counterparts).  Forward: q = phi W_a^T + b_a, or the dueling combine q = v + adv - mean(adv) with
// v = phi W_v^T + b_v.  Backward: dphi = geff W, dW += geff^T phi, db += sum geff, where geff is dq mapped through the
// dueling combine.  phi is the bf16 feature vector of the fused body; weights and gradients are fp32 (master).
// sm_100a only.
#include "common.cuh"

namespace b2rl {

constexpr int HEAD_MAX_OUT = 32;       // A (+1 for the dueling value row)

// one warp per batch row; the row of phi is read once (16-byte loads, 8 features per lane per 256-feature chunk) and
// every output accumulates against it, so all weight loads of a chunk are independent and in flight together
template <int NB>
__global__ void __launch_bounds__(128) head_fwd_kernel(const __nv_bfloat16* __restrict__ phi, const float* __restrict__ Wa,
                                                       const float* __restrict__ ba, const float* __restrict__ Wv,
                                                       const float* __restrict__ bv, int B, int K, int A,
                                                       float* __restrict__ q) {
  pdl_sync();   // PDL contract (common.cuh): before any global-memory access or return
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const __nv_bfloat16* x = phi + (int64_t)b * K;
  const int n_out = A + (Wv ? 1 : 0);
  float acc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n] = 0.0f;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const int4 xr = *reinterpret_cast<const int4*>(x + k0);
    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xr);
    float xf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(xp[j]);
      xf[2 * j] = f.x, xf[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      if (n < n_out) {
        const float* w = ((n < A) ? Wa + (int64_t)n * K : Wv) + k0;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
        float s = acc[n];
        s = fmaf(xf[0], w0.x, s), s = fmaf(xf[1], w0.y, s), s = fmaf(xf[2], w0.z, s), s = fmaf(xf[3], w0.w, s);
        s = fmaf(xf[4], w1.x, s), s = fmaf(xf[5], w1.y, s), s = fmaf(xf[6], w1.z, s), s = fmaf(xf[7], w1.w, s);
        acc[n] = s;
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    if (n < n_out) {
      const float s = warp_reduce(acc[n], OpAdd());
      acc[n] = s + ((n < A) ? ba[n] : bv[0]);
    }
  }
  if (lane == 0) {
    if (Wv) {                                     // network_heads.py:34-36: q = value + (adv - adv.mean(1))
      float mean = 0.0f, value = 0.0f;
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        if (n < A) mean += acc[n];
        if (n == A) value = acc[n];
      }
      mean /= (float)A;
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (n < A) q[(int64_t)b * A + n] = value + (acc[n] - mean);
    } else {
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (n < A) q[(int64_t)b * A + n] = acc[n];
    }
  }
}

// grid (K/64, ceil(B/HB_ROWS)); block 256 = 64 columns x 4 row groups of HB_ROWS/4 rows (many small CTAs: the kernel is
// latency-bound, 16 rows per CTA puts 256 CTAs in flight at B = 512)
constexpr int HB_ROWS = 16;
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ gq, const __nv_bfloat16* __restrict__ phi,
                                                       const float* __restrict__ Wa, const float* __restrict__ Wv, int B,
                                                       int K, int A, __nv_bfloat16* __restrict__ gphi,
                                                       float* __restrict__ gWa, float* __restrict__ gba,
                                                       float* __restrict__ gWv, float* __restrict__ gbv,
                                                       float* __restrict__ relu_colsum) {
  // relu_colsum != NULL: phi is the output of a ReLU layer (NatureConvBody's fc4): the gradient is masked here (gphi = 0 where
  // phi <= 0) and its column sums -- that layer's bias gradient -- are accumulated into relu_colsum[K] (zeroed by the caller),
  // which replaces the separate mask / bias-gradient pass over gphi.
  pdl_sync();   // PDL contract (common.cuh): before any global-memory access or return
  __shared__ float geff[HB_ROWS][HEAD_MAX_OUT + 1];       // [row][a], last used column = value gradient (dueling)
  __shared__ float red[4][64][HEAD_MAX_OUT + 1];
  const int k = blockIdx.x * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
  const int r0 = blockIdx.y * HB_ROWS;
  const int n_out = A + (Wv ? 1 : 0);
  for (int e = threadIdx.x; e < HB_ROWS * n_out; e += blockDim.x) {
    const int r = e / n_out, n = e - r * n_out;
    float g = 0.0f;
    if (r0 + r < B) {
      const float* gr = gq + (int64_t)(r0 + r) * A;
      if (!Wv) g = gr[n];
      else {
        float sum = 0.0f;
        for (int a = 0; a < A; ++a) sum += gr[a];
        g = (n < A) ? gr[n] - sum / (float)A : sum;   // d/d adv_n and d/d value of q = v + adv - mean(adv)
      }
    }
    geff[r][n] = g;
  }
  __syncthreads();
  float w[HEAD_MAX_OUT + 1], acc[HEAD_MAX_OUT + 1];
#pragma unroll
  for (int n = 0; n < HEAD_MAX_OUT + 1; ++n) {
    acc[n] = 0.0f;
    w[n] = (n < n_out && k < K) ? ((n < A) ? Wa[(int64_t)n * K + k] : Wv[k]) : 0.0f;
  }
  float colsum = 0.0f;
  if (k < K) {
    float xs[HB_ROWS / 4];
#pragma unroll
    for (int rr = 0; rr < HB_ROWS / 4; ++rr) {         // all loads in flight before the FMAs
      const int r = r0 + rg * (HB_ROWS / 4) + rr;
      xs[rr] = r < B ? __bfloat162float(phi[(int64_t)r * K + k]) : 0.0f;
    }
#pragma unroll
    for (int rr = 0; rr < HB_ROWS / 4; ++rr) {
      const int r = rg * (HB_ROWS / 4) + rr;
      if (r0 + r >= B) break;
      const float x = xs[rr];
      float g = 0.0f;
#pragma unroll
      for (int n = 0; n < HEAD_MAX_OUT + 1; ++n) {
        if (n < n_out) {
          const float ge = geff[r][n];
          g = fmaf(ge, w[n], g);
          acc[n] = fmaf(ge, x, acc[n]);
        }
      }
      if (relu_colsum && !(x > 0.0f)) g = 0.0f;
      const __nv_bfloat16 gb = __float2bfloat16_rn(g);
      colsum += __bfloat162float(gb);                  // the sum of the stored (bf16) values, as the separate pass computes it
      gphi[(int64_t)(r0 + r) * K + k] = gb;
    }
  }
  red[rg][threadIdx.x & 63][HEAD_MAX_OUT] = colsum;     // slot HEAD_MAX_OUT is free: n_out <= HEAD_MAX_OUT
#pragma unroll
  for (int n = 0; n < HEAD_MAX_OUT + 1; ++n)
    if (n < n_out) red[rg][threadIdx.x & 63][n] = acc[n];
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * n_out; e += blockDim.x) {
    const int c = e % 64, n = e / 64;
    const int kk = blockIdx.x * 64 + c;
    if (kk < K) {
      const float s = red[0][c][n] + red[1][c][n] + red[2][c][n] + red[3][c][n];
      atomicAdd((n < A) ? gWa + (int64_t)n * K + kk : gWv + kk, s);
    }
  }
  if (relu_colsum && threadIdx.x < 64) {
    const int kk = blockIdx.x * 64 + threadIdx.x;
    if (kk < K)
      atomicAdd(relu_colsum + kk, red[0][threadIdx.x][HEAD_MAX_OUT] + red[1][threadIdx.x][HEAD_MAX_OUT] +
                                      red[2][threadIdx.x][HEAD_MAX_OUT] + red[3][threadIdx.x][HEAD_MAX_OUT]);
  }
  if (blockIdx.x == 0 && threadIdx.x < n_out) {      // bias gradients: sum of geff over this block's rows
    float s = 0.0f;
    for (int r = 0; r < HB_ROWS; ++r) s += geff[r][threadIdx.x];
    atomicAdd((threadIdx.x < A) ? gba + threadIdx.x : gbv, s);
  }
}

}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_head_fwd(const uint16_t* phi, const float* Wa, const float* ba, const float* Wv, const float* bv,
                             int32_t B, int32_t K, int32_t A, float* q, void* stream) {
  B2RL_REQUIRE(phi && Wa && ba && q && ((Wv == nullptr) == (bv == nullptr)), "null pointer");
  B2RL_REQUIRE(B > 0 && K > 0 && K % 8 == 0 && A > 0 && A < HEAD_MAX_OUT, "need K % 8 == 0 and 0 < A < 32");
  B2RL_REQUIRE(reinterpret_cast<uintptr_t>(phi) % 16 == 0 && (reinterpret_cast<uintptr_t>(Wa) | reinterpret_cast<uintptr_t>(Wv)) % 16 == 0,
               "phi and the weights must be 16-byte aligned");
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(phi);
  const int n_out = A + (Wv ? 1 : 0);
  const dim3 grid((B + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_out <= 8) launch_pdl(head_fwd_kernel<8>, dim3(grid), dim3(128), 0, st, x, Wa, ba, Wv, bv, B, K, A, q);
  else if (n_out <= 19) launch_pdl(head_fwd_kernel<19>, dim3(grid), dim3(128), 0, st, x, Wa, ba, Wv, bv, B, K, A, q);
  else launch_pdl(head_fwd_kernel<HEAD_MAX_OUT>, dim3(grid), dim3(128), 0, st, x, Wa, ba, Wv, bv, B, K, A, q);
  return check_launch("b2rl_head_fwd");
}

static int head_bwd_impl(const float* gq, const uint16_t* phi, const float* Wa, const float* Wv, int32_t B, int32_t K,
                         int32_t A, uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, float* relu_colsum,
                         void* stream) {
  B2RL_REQUIRE(gq && phi && Wa && gphi && gWa && gba && ((Wv == nullptr) == (gWv == nullptr)) && ((Wv == nullptr) == (gbv == nullptr)),
               "null pointer");
  B2RL_REQUIRE(B > 0 && K > 0 && A > 0 && A < HEAD_MAX_OUT, "need 0 < A < 32");
  dim3 grid((K + 63) / 64, (B + HB_ROWS - 1) / HB_ROWS);
  launch_pdl(head_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, gq, reinterpret_cast<const __nv_bfloat16*>(phi), Wa, Wv, B, K, A,
                                                          reinterpret_cast<__nv_bfloat16*>(gphi), gWa, gba, gWv, gbv, relu_colsum);
  return check_launch("b2rl_head_bwd");
}

extern "C" int b2rl_head_bwd(const float* gq, const uint16_t* phi, const float* Wa, const float* Wv, int32_t B, int32_t K,
                             int32_t A, uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, void* stream) {
  return head_bwd_impl(gq, phi, Wa, Wv, B, K, A, gphi, gWa, gba, gWv, gbv, nullptr, stream);
}

// b2rl_head_bwd for features phi = relu(layer(.)): gphi is masked (0 where phi <= 0) and relu_colsum[K] += column sums of the
// masked gphi (the bias gradient of that layer; zero it first) -- the ReLU backward of the body's last layer in the same pass
extern "C" int b2rl_head_bwd_relu(const float* gq, const uint16_t* phi, const float* Wa, const float* Wv, int32_t B, int32_t K,
                                  int32_t A, uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv,
                                  float* relu_colsum, void* stream) {
  B2RL_REQUIRE(relu_colsum, "null relu_colsum");
  return head_bwd_impl(gq, phi, Wa, Wv, B, K, A, gphi, gWa, gba, gWv, gbv, relu_colsum, stream);
}
