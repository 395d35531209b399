// Attention cores (head dim 64) for the Hi3D VideoUNet.
//   fmha_d64_kernel      spatial self-attention, L = H*W keys per image  (flash-style online softmax)
//   tattn_d64_kernel     temporal self-attention over T <= 16 frames per (clip, pixel, head)
// Both read a packed fp16 [rows, 3C] q|k|v matrix produced by one fused QKV GEMM and write fp16 [rows, C].
#include <string.h>

#include "common.cuh"

namespace hi3d {

// ================================================================================================
// Spatial FMHA: CTA = 128 queries x one (image, head); 8 warps x 16 query rows; KV tiles of 64.
// ================================================================================================
constexpr int FQ = 128;   // queries per CTA
constexpr int FK = 64;    // keys per tile
constexpr int FMHA_THREADS = 256;
constexpr int FMHA_SMEM = FQ * 128 + 2 * FK * 128 + 2 * FK * 128;  // Q + 2xK + 2xV = 48 KB

__global__ void __launch_bounds__(FMHA_THREADS, 2)
fmha_d64_kernel(const __half* __restrict__ qkv, int L, int C, float scale_log2, __half* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + FQ * 128;
  const uint32_t sV = sK + 2 * FK * 128;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * FQ;
  const int h = blockIdx.y;
  const long long tok0 = (long long)blockIdx.z * L;
  const int ld = 3 * C;
  const __half* qbase = qkv + tok0 * ld + h * 64;
  const __half* kbase = qbase + C;
  const __half* vbase = qbase + 2 * C;
  const int chunk = tid & 7, rbase = tid >> 3;  // rbase 0..31

  // ---- issue Q + first K/V tile ----
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = rbase + 32 * i;
    const bool v = (q0 + r) < L;
    cp_async16(sQ + swz128(r, chunk), v ? qbase + (long long)(q0 + r) * ld + chunk * 8 : qbase, v);
  }
  auto load_kv = [&](int j, int buf) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int r = rbase + 32 * i;
      const int kv = j * FK + r;
      const bool v = kv < L;
      cp_async16(sK + buf * (FK * 128) + swz128(r, chunk), v ? kbase + (long long)kv * ld + chunk * 8 : kbase, v);
      cp_async16(sV + buf * (FK * 128) + swz128(r, chunk), v ? vbase + (long long)kv * ld + chunk * 8 : vbase, v);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int nkv = (L + FK - 1) / FK;
  uint32_t aq[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int k = 0; k < 4; k++) o[i][k] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const int g = lane >> 2, t4 = lane & 3;

  for (int j = 0; j < nkv; j++) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_kv(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int k16 = 0; k16 < 4; k16++)
        ldmatrix_x4(aq[k16][0], aq[k16][1], aq[k16][2], aq[k16][3],
                    sQ + swz128(warp * 16 + (lane & 15), k16 * 2 + (lane >> 4)));
    }
    // ---- S = Q K^T ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int k = 0; k < 4; k++) s[i][k] = 0.f;
    const uint32_t sKb = sK + buf * (FK * 128);
#pragma unroll
    for (int k16 = 0; k16 < 4; k16++) {
#pragma unroll
      for (int np = 0; np < 4; np++) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, sKb + swz128(np * 16 + (lane & 7) + ((lane >> 4) << 3), k16 * 2 + ((lane >> 3) & 1)));
        mma_16816(s[2 * np], aq[k16], b0, b1);
        mma_16816(s[2 * np + 1], aq[k16], b2, b3);
      }
    }
    // ---- mask the key tail ----
    if ((j + 1) * FK > L) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const int col = j * FK + nt * 8 + 2 * t4;
        if (col >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (col + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    // ---- online softmax ----
    uint32_t pa[4][4];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) mx = fmaxf(mx, fmaxf(s[nt][2 * hh], s[nt][2 * hh + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float mnew = fmaxf(mrow[hh], mx);
      const float corr = exp2f((mrow[hh] - mnew) * scale_log2);
      const float moff = mnew * scale_log2;
      float rs = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const float p0 = exp2f(s[nt][2 * hh] * scale_log2 - moff);
        const float p1 = exp2f(s[nt][2 * hh + 1] * scale_log2 - moff);
        rs += p0 + p1;
        // C-fragment -> A-fragment: key tile nt feeds k16 step nt/2, register (nt&1)*2 + hh
        pa[nt >> 1][(nt & 1) * 2 + hh] = pack_half2(p0, p1);
      }
      lrow[hh] = lrow[hh] * corr + rs;
      mrow[hh] = mnew;
#pragma unroll
      for (int dt = 0; dt < 8; dt++) {
        o[dt][2 * hh] *= corr;
        o[dt][2 * hh + 1] *= corr;
      }
    }
    // ---- O += P V ----
    const uint32_t sVb = sV + buf * (FK * 128);
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int dp = 0; dp < 4; dp++) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3,
                          sVb + swz128(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
        mma_16816(o[2 * dp], pa[kk], b0, b1);
        mma_16816(o[2 * dp + 1], pa[kk], b2, b3);
      }
    }
    __syncthreads();
  }

  // ---- finalize: O /= l, stage through this warp's own Q rows, 16-byte stores ----
#pragma unroll
  for (int hh = 0; hh < 2; hh++) {
    float l = lrow[hh];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.f / l;
    const int r = warp * 16 + g + 8 * hh;
#pragma unroll
    for (int dt = 0; dt < 8; dt++) {
      const uint32_t v = pack_half2(o[dt][2 * hh] * inv, o[dt][2 * hh + 1] * inv);
      // logical position: row r, 16B chunk dt, element offset 2*t4 inside the chunk
      *reinterpret_cast<uint32_t*>(smem + swz128(r, dt) + t4 * 4) = v;
    }
  }
  __syncwarp();
  __half* obase = out + tok0 * C + h * 64;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = warp * 16 + (lane >> 3) + 4 * i;
    const int c = lane & 7;
    if (q0 + r < L) {
      const uint4 v = *reinterpret_cast<const uint4*>(smem + swz128(r, c));
      *reinterpret_cast<uint4*>(obase + (long long)(q0 + r) * C + c * 8) = v;
    }
  }
}

// ================================================================================================
// Temporal attention: one warp per (clip b, pixel s, head h); T <= 16 frames.
// ================================================================================================
constexpr int TA_WARPS = 4;

// One launch covers the pixel strip [s0, s0 + s_cnt) of every clip for ALL T = Tl * world frames.  Unsharded runs have
// world = 1 (one buffer, the whole pixel range).  Frame-sharded runs: rank r owns frames [r Tl, (r+1) Tl) of
// every clip in ITS qkv / out buffers and computes strip r of the pixels -- the q|k|v rows of the other ranks' frames are
// read straight from their (IPC-mapped) buffers over NVLink and the output rows of their frames are stored straight into
// their buffers: the "all-gather before temporal attention" never materialises, no rank computes a row twice, and 1/4 of the
// bytes of a K/V gather cross the links (q|k|v of 1/world of the pixels in, 1/world of the outputs out).
struct TaParams {
  const __half* qkv[HI3D_MAX_PEERS];   // [B * Tl * S, 3C] of the rank that owns frames [r Tl, (r+1) Tl)
  __half* out[HI3D_MAX_PEERS];         // [B * Tl * S, C]
  int B, Tl, world, S, heads;
  int s0, s_cnt;
  float scale_log2;
  long long n_items;                   // B * s_cnt * heads
};

__global__ void __launch_bounds__(TA_WARPS * 32) tattn_d64_kernel(const __grid_constant__ TaParams p) {
  __shared__ __align__(1024) uint8_t smem[TA_WARPS * 3 * 2048];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long item = (long long)blockIdx.x * TA_WARPS + warp;
  if (item >= p.n_items) return;
  const int heads = p.heads, S = p.S, Tl = p.Tl, T = p.Tl * p.world;
  const float scale_log2 = p.scale_log2;
  const int C = heads * 64;
  const int h = (int)(item % heads);
  const long long bs = item / heads;
  const int s = p.s0 + (int)(bs % p.s_cnt);
  const int b = (int)(bs / p.s_cnt);
  const int ld = 3 * C;
  const uint32_t sQ = smem_u32(smem) + warp * (3 * 2048);
  const uint32_t sK = sQ + 2048, sV = sQ + 4096;
  uint8_t* sQg = smem + warp * (3 * 2048);

  const int chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int t = (lane >> 3) + 4 * i;
    const bool v = t < T;
    const int tt = v ? t : 0;
    const int owner = tt / Tl, tl = tt - owner * Tl;
    const __half* row = p.qkv[owner] + ((long long)(b * Tl + tl) * S + s) * ld + h * 64 + chunk * 8;
    cp_async16(sQ + swz128(t, chunk), row, v);
    cp_async16(sK + swz128(t, chunk), row + C, v);
    cp_async16(sV + swz128(t, chunk), row + 2 * C, v);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  float sc[2][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int k = 0; k < 4; k++) sc[i][k] = 0.f;
#pragma unroll
  for (int k16 = 0; k16 < 4; k16++) {
    uint32_t a[4], b0, b1, b2, b3;
    ldmatrix_x4(a[0], a[1], a[2], a[3], sQ + swz128(lane & 15, k16 * 2 + (lane >> 4)));
    ldmatrix_x4(b0, b1, b2, b3, sK + swz128((lane & 7) + ((lane >> 4) << 3), k16 * 2 + ((lane >> 3) & 1)));
    mma_16816(sc[0], a, b0, b1);
    mma_16816(sc[1], a, b2, b3);
  }
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; nt++) {
    const int col = nt * 8 + 2 * t4;
    if (col >= T) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
    if (col + 1 >= T) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
  }
  uint32_t pa[4];
  float inv[2];
#pragma unroll
  for (int hh = 0; hh < 2; hh++) {
    float mx = fmaxf(fmaxf(sc[0][2 * hh], sc[0][2 * hh + 1]), fmaxf(sc[1][2 * hh], sc[1][2 * hh + 1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float moff = mx * scale_log2;
    const float p00 = exp2f(sc[0][2 * hh] * scale_log2 - moff), p01 = exp2f(sc[0][2 * hh + 1] * scale_log2 - moff);
    const float p10 = exp2f(sc[1][2 * hh] * scale_log2 - moff), p11 = exp2f(sc[1][2 * hh + 1] * scale_log2 - moff);
    float rs = p00 + p01 + p10 + p11;
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    inv[hh] = 1.f / rs;
    pa[hh] = pack_half2(p00, p01);       // a0 (row g) / a1 (row g+8): keys 2t..2t+1
    pa[2 + hh] = pack_half2(p10, p11);   // a2 / a3: keys 8+2t..
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int k = 0; k < 4; k++) o[i][k] = 0.f;
#pragma unroll
  for (int dp = 0; dp < 4; dp++) {
    uint32_t b0, b1, b2, b3;
    ldmatrix_x4_trans(b0, b1, b2, b3, sV + swz128((lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
    mma_16816(o[2 * dp], pa, b0, b1);
    mma_16816(o[2 * dp + 1], pa, b2, b3);
  }
  __syncwarp();   // all lanes are done reading sQ (ldmatrix) before it is reused as the output stage
#pragma unroll
  for (int hh = 0; hh < 2; hh++) {
    const int r = g + 8 * hh;
#pragma unroll
    for (int dt = 0; dt < 8; dt++)
      *reinterpret_cast<uint32_t*>(sQg + swz128(r, dt) + t4 * 4) =
          pack_half2(o[dt][2 * hh] * inv[hh], o[dt][2 * hh + 1] * inv[hh]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int t = (lane >> 3) + 4 * i;
    if (t < T) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQg + swz128(t, chunk));
      const int owner = t / Tl, tl = t - owner * Tl;
      *reinterpret_cast<uint4*>(p.out[owner] + ((long long)(b * Tl + tl) * S + s) * C + h * 64 + chunk * 8) = v;
    }
  }
  if (p.world > 1) __threadfence_system();     // peer stores of the output rows are performed before the kernel ends
}

// ================================================================================================
// Row softmax (in place, fp16) and transpose: building blocks for the VAE d=512 single-head attention.
// ================================================================================================
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ s, int L, float scale_log2) {
  __shared__ float red[8];
  __half* row = s + (long long)blockIdx.x * L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int i = tid * 8; i < L; i += 256 * 8) {
    Half8 v = *reinterpret_cast<const Half8*>(row + i);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float2 f = __half22float2(v.h[q]);
      mx = fmaxf(mx, fmaxf(f.x, f.y));
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float moff = mx * scale_log2;
  float sum = 0.f;
  for (int i = tid * 8; i < L; i += 256 * 8) {
    Half8 v = *reinterpret_cast<const Half8*>(row + i);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float2 f = __half22float2(v.h[q]);
      sum += exp2f(f.x * scale_log2 - moff) + exp2f(f.y * scale_log2 - moff);
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; w++) sum += red[w];
  const float inv = 1.f / sum;
  for (int i = tid * 8; i < L; i += 256 * 8) {
    Half8 v = *reinterpret_cast<const Half8*>(row + i);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float2 f = __half22float2(v.h[q]);
      v.h[q] = __floats2half2_rn(exp2f(f.x * scale_log2 - moff) * inv, exp2f(f.y * scale_log2 - moff) * inv);
    }
    *reinterpret_cast<Half8*>(row + i) = v;
  }
}

__global__ void __launch_bounds__(256) transpose_kernel(const __half* __restrict__ in, int R, int Cc, int in_ld,
                                                        __half* __restrict__ out) {
  __shared__ __half tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < Cc) ? in[(long long)r * in_ld + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < Cc && r < R) out[(long long)c * R + r] = tile[tx][i];
  }
}

}  // namespace hi3d

using namespace hi3d;

extern "C" int hi3d_attention_d64(const void* qkv, int n_img, int L, int heads, float scale, void* out, void* stream) {
  if (!qkv || !out || n_img <= 0 || L <= 0 || heads <= 0 || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) {
    set_error("hi3d_attention_d64: bad arguments (n_img=%d L=%d heads=%d)", n_img, L, heads);
    return -2;
  }
  if (heads > 65535 || n_img > 65535) { set_error("hi3d_attention_d64: grid too large"); return -2; }
  static bool attr_done[HI3D_MAX_DEVICES];
  if (ensure_dyn_smem(fmha_d64_kernel, FMHA_SMEM, attr_done, "hi3d_attention_d64")) return -1;
  dim3 grid((L + FQ - 1) / FQ, heads, n_img);
  fmha_d64_kernel<<<grid, FMHA_THREADS, FMHA_SMEM, (cudaStream_t)stream>>>(
      (const __half*)qkv, L, heads * 64, scale * 1.4426950408889634f, (__half*)out);
  return check_launch("hi3d_attention_d64");
}

static int launch_tattn(const TaParams& tp, cudaStream_t st, const char* who) {
  const long long blocks = (tp.n_items + TA_WARPS - 1) / TA_WARPS;
  if (blocks > 2147483647LL) { set_error("%s: too many items", who); return -2; }
  if (blocks > 0) tattn_d64_kernel<<<(unsigned)blocks, TA_WARPS * 32, 0, st>>>(tp);
  return check_launch(who);
}

extern "C" int hi3d_temporal_attention_d64(const void* qkv, int B, int T, int S, int heads, float scale, void* out,
                                           void* stream) {
  if (!qkv || !out || B <= 0 || T <= 0 || T > 16 || S <= 0 || heads <= 0 || ((uintptr_t)qkv & 15) ||
      ((uintptr_t)out & 15)) {
    set_error("hi3d_temporal_attention_d64: bad arguments (B=%d T=%d S=%d heads=%d); T must be <= 16", B, T, S, heads);
    return -2;
  }
  TaParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.qkv[0] = (const __half*)qkv; tp.out[0] = (__half*)out;
  tp.B = B; tp.Tl = T; tp.world = 1; tp.S = S; tp.heads = heads; tp.s0 = 0; tp.s_cnt = S;
  tp.scale_log2 = scale * 1.4426950408889634f;
  tp.n_items = (long long)B * S * heads;
  return launch_tattn(tp, (cudaStream_t)stream, "hi3d_temporal_attention_d64");
}

extern "C" int hi3d_temporal_attention_d64_sharded(void* const* qkv_of_rank, void* const* out_of_rank, int rank, int world,
                                                   int B, int T_local, int S, int heads, float scale, void* stream) {
  if (!qkv_of_rank || !out_of_rank || world < 1 || world > HI3D_MAX_PEERS || rank < 0 || rank >= world || B <= 0 ||
      T_local <= 0 || T_local * world > 16 || S <= 0 || heads <= 0) {
    set_error("hi3d_temporal_attention_d64_sharded: bad arguments (rank=%d world=%d B=%d T_local=%d S=%d heads=%d); "
              "T_local * world must be <= 16", rank, world, B, T_local, S, heads);
    return -2;
  }
  TaParams tp;
  memset(&tp, 0, sizeof(tp));
  for (int r = 0; r < world; r++) {
    if (!qkv_of_rank[r] || !out_of_rank[r] || ((uintptr_t)qkv_of_rank[r] & 15) || ((uintptr_t)out_of_rank[r] & 15)) {
      set_error("hi3d_temporal_attention_d64_sharded: null / unaligned buffer of rank %d", r);
      return -2;
    }
    tp.qkv[r] = (const __half*)qkv_of_rank[r]; tp.out[r] = (__half*)out_of_rank[r];
  }
  const int per = (S + world - 1) / world;             // pixel strip of this rank (the last strips may be short / empty)
  int s0 = rank * per, s1 = s0 + per;
  if (s0 > S) s0 = S;
  if (s1 > S) s1 = S;
  tp.B = B; tp.Tl = T_local; tp.world = world; tp.S = S; tp.heads = heads; tp.s0 = s0; tp.s_cnt = s1 - s0;
  tp.scale_log2 = scale * 1.4426950408889634f;
  tp.n_items = (long long)B * tp.s_cnt * heads;
  return launch_tattn(tp, (cudaStream_t)stream, "hi3d_temporal_attention_d64_sharded");
}

extern "C" int hi3d_softmax_rows(void* s, int64_t rows, int L, float scale, void* stream) {
  if (!s || rows <= 0 || L <= 0 || (L % 8) || ((uintptr_t)s & 15) || rows > 2147483647LL) {
    set_error("hi3d_softmax_rows: bad arguments (rows=%lld L=%d; L must be a multiple of 8)", (long long)rows, L);
    return -2;
  }
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((__half*)s, L, scale * 1.4426950408889634f);
  return check_launch("hi3d_softmax_rows");
}

extern "C" int hi3d_transpose(const void* in, int R, int Cc, int in_ld, void* out, void* stream) {
  if (!in || !out || R <= 0 || Cc <= 0 || in_ld < Cc) { set_error("hi3d_transpose: bad arguments"); return -2; }
  dim3 grid((Cc + 31) / 32, (R + 31) / 32);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)in, R, Cc, in_ld, (__half*)out);
  return check_launch("hi3d_transpose");
}
