// Implicit-GEMM engine, Blackwell-native variant: TMA (cp.async.bulk.tensor) operand staging into
// SWIZZLE_128B shared tiles, tcgen05.mma (UTCHMMA, cta_group::1, M=128) with the fp32 accumulator in TMEM,
// tcgen05.ld epilogue.  Same contract as hi3d_gemm (include/hi3d_b200.h); geometries this engine does not
// cover (stride-2 / upsample-fused convs, odd tile shapes) are forwarded to the mma.sync engine.
//
// One CTA = one 128 x BN output tile.  The 128 rows of a tile are
//   PLAIN    : 128 consecutive rows
//   CONV2D   : a (tn images) x (th rows) x (tw columns) patch, tn*th*tw = 128 -- the 3x3 taps then are the same
//              TMA box shifted by (dy, dx), and the zero padding is TMA out-of-bounds fill
//   TEMPORAL : (tf frames) x (ts pixels), tf*ts = 128 -- the temporal taps shift the frame coordinate, clip
//              boundaries zero-fill by OOB on the frame axis of a [C, HW, T, B] view.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace hi3d {
int validate_gemm(const hi3d_gemm_params* p, const char* who);

constexpr int T5_BM = 128;
constexpr int T5_BK = 64;
constexpr int T5_THREADS = 192;
constexpr int T5_MAX_MAPS = 4;

struct T5Seg {
  int map;       // index into amap[]
  int c_off, C;  // channel range
  int dy, dx, dt;
};

struct T5Params {
  CUtensorMap bmap;
  CUtensorMap amap[T5_MAX_MAPS];
  T5Seg seg[HI3D_MAX_SEGS];
  int nseg;
  int M, N, K, mode;
  // tile -> rows
  int tw, th, tn;      // CONV2D patch (PLAIN: tw = 128, th = tn = 1; TEMPORAL: tw = ts, th = tf)
  int Wo, Ho, Nimg;    // CONV2D: output W, H, images.  TEMPORAL: Wo = HW, Ho = T, Nimg = B
  int tiles_x, tiles_y;  // tiles along W and H (CONV2D) / along HW and T (TEMPORAL)
  // epilogue
  const float* bias;
  const __half* rowbias;
  int rb_div, rb_mod, rb_ld, act;
  const __half* residual;
  int res_ld;
  const __half* blend_x;
  int blend_ld;
  float alpha;
  __half* out;
  int out_ld;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------
HI3D_DEVINL void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
HI3D_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
HI3D_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1ull << 26)) __trap();  // watchdog: a lost TMA / MMA completion must not hang the GPU
  }
}
HI3D_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
HI3D_DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
HI3D_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
HI3D_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
HI3D_DEVINL void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
HI3D_DEVINL void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// SWIZZLE_128B, K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 |
// SBO(1024B>>4)<<32 | version 1 @46 | layout SWIZZLE_128B (2) @61.
HI3D_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
HI3D_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

template <int BN, int STAGES>
struct T5Smem {
  static constexpr int A_BYTES = T5_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = T5_BM * (BN + 8) * 2;
  static constexpr int DATA_BYTES = PIPE_BYTES > STAGING_BYTES ? PIPE_BYTES : STAGING_BYTES;
  static constexpr int BAR_OFF = DATA_BYTES;              // full[STAGES], empty[STAGES], accum, tmem ptr
  static constexpr int TOTAL = DATA_BYTES + 128 + 1024;   // + barriers + 1 KB alignment slack
};

// tile-local row -> global output row (and validity)
HI3D_DEVINL long long t5_row(const T5Params& p, int tile, int r) {
  if (p.mode == HI3D_ROWS_PLAIN) {
    long long m = (long long)tile * T5_BM + r;
    return m < p.M ? m : -1;
  }
  const int tx = tile % p.tiles_x;
  const int rest = tile / p.tiles_x;
  const int ty = rest % p.tiles_y;
  const int tz = rest / p.tiles_y;
  const int x = tx * p.tw + r % p.tw;
  const int y = ty * p.th + (r / p.tw) % p.th;
  const int n = tz * p.tn + r / (p.tw * p.th);
  if (x >= p.Wo || y >= p.Ho || n >= p.Nimg) return -1;
  return ((long long)n * p.Ho + y) * p.Wo + x;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(T5_THREADS, 1) gemm_tc5_kernel(const __grid_constant__ T5Params p) {
  using SM = T5Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B atoms
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_full = base + SM::BAR_OFF;            // STAGES x 8 bytes
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_accum = bar_empty + 8 * STAGES;
  const uint32_t tmem_slot = bar_accum + 8;
  volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(smem + SM::BAR_OFF + 16 * STAGES + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int KT = p.K / T5_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      // tile origin coordinates
      int x0 = 0, y0 = 0, z0 = 0;
      if (p.mode != HI3D_ROWS_PLAIN) {
        const int tx = tile % p.tiles_x, rest = tile / p.tiles_x;
        x0 = tx * p.tw;
        y0 = (rest % p.tiles_y) * p.th;
        z0 = (rest / p.tiles_y) * p.tn;
      }
      int si = 0, so = 0;
      for (int kt = 0; kt < KT; kt++) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t sA = base + s * SM::STAGE_BYTES;
        const uint32_t sB = sA + SM::A_BYTES;
        mbar_expect_tx(bar_full + 8 * s, SM::STAGE_BYTES);
        const T5Seg sg = p.seg[si];
        const int c = sg.c_off + so;
        if (p.mode == HI3D_ROWS_PLAIN)
          tma_load_2d(sA, &p.amap[sg.map], bar_full + 8 * s, c, tile * T5_BM);
        else if (p.mode == HI3D_ROWS_CONV2D)
          tma_load_4d(sA, &p.amap[sg.map], bar_full + 8 * s, c, x0 + sg.dx, y0 + sg.dy, z0);
        else
          tma_load_4d(sA, &p.amap[sg.map], bar_full + 8 * s, c, x0, y0 + sg.dt, z0);
        tma_load_2d(sB, &p.bmap, bar_full + 8 * s, kt * T5_BK, n0);
        so += T5_BK;
        if (so >= sg.C) { si++; so = 0; }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T5_BM >> 4) << 24);
      for (int kt = 0; kt < KT; kt++) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const uint32_t sA = base + s * SM::STAGE_BYTES;
        const uint32_t sB = sA + SM::A_BYTES;
        const uint64_t ad = umma_desc_sw128(sA), bd = umma_desc_sw128(sB);
#pragma unroll
        for (int k = 0; k < T5_BK / 16; k++)   // +32 bytes (2 x 16 B) along K inside the 128-byte swizzle atom
          tc_mma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kt | k) ? 1u : 0u);
        tc_commit(bar_empty + 8 * s);            // frees the smem slot when these MMAs retire
      }
      tc_commit(bar_accum);                      // accumulator complete
    }
  } else {
    // ======================= epilogue warps (2..5) =======================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    mbar_wait(bar_accum, 0);
    tc_fence_after();
    const bool geglu = (p.act == HI3D_ACT_GEGLU);
    const int pitch = BN + 8;
    __half* sC = reinterpret_cast<__half*>(smem);
    const int rl = q * 32 + lane;                // tile-local row == TMEM lane
    const long long m = t5_row(p, tile, rl);
    const __half* rbp = nullptr;
    if (p.rowbias != nullptr && m >= 0) rbp = p.rowbias + (long long)((m / p.rb_div) % p.rb_mod) * p.rb_ld;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const int n = n0 + c0 + j;
        float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
        if (n < p.N) {
          if (p.bias != nullptr) { a += __ldg(p.bias + n); b += __ldg(p.bias + n + 1); }
          if (rbp != nullptr) {
            const __half2 rb = *reinterpret_cast<const __half2*>(rbp + n);
            a += __low2float(rb); b += __high2float(rb);
          }
        }
        if (geglu) {
          sC[rl * pitch + ((c0 + j) >> 1)] = __float2half_rn(a * gelu_erf_f(b));
        } else {
          if (p.act == HI3D_ACT_SILU) { a = silu_f(a); b = silu_f(b); }
          *reinterpret_cast<uint32_t*>(sC + rl * pitch + c0 + j) = pack_half2(a, b);
        }
      }
    }
    tc_fence_before();
    // the 4 epilogue warps synchronise among themselves (named barrier 1, 128 threads)
    asm volatile("bar.sync 1, 128;\n" ::: "memory");
    const int et = tid - 64;                     // 0..127
    const int BNo = geglu ? BN / 2 : BN;
    const int cpr = BNo / 8;
    const int Nout = geglu ? p.N / 2 : p.N;
    const int nout0 = geglu ? n0 / 2 : n0;
    for (int idx = et; idx < T5_BM * cpr; idx += 128) {
      const int r = idx / cpr, c = idx - r * cpr;
      const int nc = nout0 + c * 8;
      const long long mm = t5_row(p, tile, r);
      if (mm < 0 || nc >= Nout) continue;
      Half8 v = *reinterpret_cast<const Half8*>(sC + r * pitch + c * 8);
      if (p.residual != nullptr) {
        const Half8 rr = *reinterpret_cast<const Half8*>(p.residual + mm * p.res_ld + nc);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float2 x = __half22float2(v.h[k]), y = __half22float2(rr.h[k]);
          v.h[k] = __floats2half2_rn(x.x + y.x, x.y + y.y);
        }
      }
      if (p.blend_x != nullptr) {
        const Half8 xx = *reinterpret_cast<const Half8*>(p.blend_x + mm * p.blend_ld + nc);
        const float al = p.alpha, be = 1.f - p.alpha;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float2 x = __half22float2(v.h[k]), y = __half22float2(xx.h[k]);
          v.h[k] = __floats2half2_rn(al * y.x + be * x.x, al * y.y + be * x.y);
        }
      }
      *reinterpret_cast<Half8*>(p.out + mm * p.out_ld + nc) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(BN));
  }
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("hi3d_gemm_tc5: cuTensorMapEncodeTiled entry point unavailable"); return -1; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("hi3d_gemm_tc5: cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return -1;
  }
  return 0;
}

static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <int BN>
static int launch_tc5(const T5Params& tp, int tiles, cudaStream_t st) {
  constexpr int STAGES = (BN == 128) ? 3 : 4;   // BN=128: 97 KB -> two CTAs per SM
  using SM = T5Smem<BN, STAGES>;
  auto kern = gemm_tc5_kernel<BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) { set_error("hi3d_gemm_tc5: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_done = true;
  }
  dim3 grid(tiles, (tp.N + BN - 1) / BN);
  kern<<<grid, T5_THREADS, SM::TOTAL, st>>>(tp);
  return check_launch("hi3d_gemm_tc5");
}

}  // namespace hi3d

using namespace hi3d;

extern "C" int hi3d_gemm_tc5(const hi3d_gemm_params* p, void* stream) {
  int rc = validate_gemm(p, "hi3d_gemm_tc5");
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // ---- geometry this engine covers; everything else goes to the mma.sync engine (same results) ----
  T5Params tp;
  memset(&tp, 0, sizeof(tp));
  tp.M = p->M; tp.N = p->N; tp.K = p->K; tp.mode = p->mode; tp.nseg = p->nseg;
  int tiles = 0;
  bool ok = (p->N >= 64);
  if (p->mode == HI3D_ROWS_PLAIN) {
    tp.tw = 128; tp.th = 1; tp.tn = 1;
    tiles = (p->M + T5_BM - 1) / T5_BM;
  } else if (p->mode == HI3D_ROWS_CONV2D) {
    ok = ok && p->stride == 1 && p->ups == 0 && p->Ho == p->Hs && p->Wo == p->Ws;
    const int Nimg = ok ? p->M / (p->Ho * p->Wo) : 0;
    int tw = 1;
    while (tw * 2 <= 16 && (p->Wo % (tw * 2)) == 0) tw *= 2;
    int th = 1;
    while (tw * th * 2 <= 128 && (p->Ho % (th * 2)) == 0) th *= 2;
    int tn = ok ? 128 / (tw * th) : 1;
    ok = ok && pow2(tn) && tw * th * tn == 128 && (Nimg % tn) == 0;
    tp.tw = tw; tp.th = th; tp.tn = tn; tp.Wo = p->Wo; tp.Ho = p->Ho; tp.Nimg = Nimg;
    tp.tiles_x = p->Wo / tw; tp.tiles_y = p->Ho / th;
    tiles = ok ? tp.tiles_x * tp.tiles_y * (Nimg / tn) : 0;
  } else {
    const int HW = p->Ho * p->Wo, T = p->T, B = p->M / (HW * T);
    int ts = 1;
    while (ts * 2 <= 128 && (HW % (ts * 2)) == 0) ts *= 2;
    int tf = 128 / ts;
    ok = ok && (T % tf) == 0;
    tp.tw = ts; tp.th = tf; tp.tn = 1; tp.Wo = HW; tp.Ho = T; tp.Nimg = B;
    tp.tiles_x = HW / ts; tp.tiles_y = ok ? T / tf : 1;
    tiles = ok ? tp.tiles_x * tp.tiles_y * B : 0;
  }
  // distinct A sources -> tensor maps
  const void* srcs[T5_MAX_MAPS];
  int lds[T5_MAX_MAPS];
  int nmaps = 0;
  for (int i = 0; ok && i < p->nseg; i++) {
    const hi3d_seg& s = p->seg[i];
    int mi = -1;
    for (int j = 0; j < nmaps; j++)
      if (srcs[j] == s.src && lds[j] == s.ld) mi = j;
    if (mi < 0) {
      if (nmaps == T5_MAX_MAPS) { ok = false; break; }
      srcs[nmaps] = s.src; lds[nmaps] = s.ld; mi = nmaps++;
    }
    tp.seg[i].map = mi; tp.seg[i].c_off = s.c_off; tp.seg[i].C = s.C;
    tp.seg[i].dy = s.dy; tp.seg[i].dx = s.dx; tp.seg[i].dt = s.dt;
  }
  if (!ok) return hi3d_gemm(p, stream);

  // tile-N: least padded work; ties go to the wider tile, the 64-wide tile pays a 10% penalty
  int BN = 256;
  {
    double best = 1e30;
    const int cand[3] = {256, 128, 64};
    for (int i = 0; i < 3; i++) {
      double c = (double)((p->N + cand[i] - 1) / cand[i]) * cand[i] * (cand[i] == 64 ? 1.1 : 1.0);
      if (c < best - 1e-9) { best = c; BN = cand[i]; }
    }
  }
  for (int j = 0; j < nmaps; j++) {
    const cuuint64_t ld = (cuuint64_t)lds[j];
    if (p->mode == HI3D_ROWS_PLAIN) {
      cuuint64_t dims[2] = {ld, (cuuint64_t)p->M};
      cuuint64_t str[1] = {ld * 2};
      cuuint32_t box[2] = {64, 128};
      if (encode_map(&tp.amap[j], srcs[j], 2, dims, str, box)) return -1;
    } else if (p->mode == HI3D_ROWS_CONV2D) {
      cuuint64_t dims[4] = {ld, (cuuint64_t)p->Ws, (cuuint64_t)p->Hs, (cuuint64_t)tp.Nimg};
      cuuint64_t str[3] = {ld * 2, ld * 2 * p->Ws, ld * 2 * p->Ws * p->Hs};
      cuuint32_t box[4] = {64, (cuuint32_t)tp.tw, (cuuint32_t)tp.th, (cuuint32_t)tp.tn};
      if (encode_map(&tp.amap[j], srcs[j], 4, dims, str, box)) return -1;
    } else {
      const cuuint64_t HW = (cuuint64_t)tp.Wo, T = (cuuint64_t)tp.Ho;
      cuuint64_t dims[4] = {ld, HW, T, (cuuint64_t)tp.Nimg};
      cuuint64_t str[3] = {ld * 2, ld * 2 * HW, ld * 2 * HW * T};
      cuuint32_t box[4] = {64, (cuuint32_t)tp.tw, (cuuint32_t)tp.th, 1};
      if (encode_map(&tp.amap[j], srcs[j], 4, dims, str, box)) return -1;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p->K, (cuuint64_t)p->N};
    cuuint64_t str[1] = {(cuuint64_t)p->K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    if (encode_map(&tp.bmap, p->W, 2, dims, str, box)) return -1;
  }
  tp.bias = p->bias; tp.rowbias = (const __half*)p->rowbias; tp.rb_div = p->rb_div; tp.rb_mod = p->rb_mod;
  tp.rb_ld = p->rb_ld; tp.act = p->act; tp.residual = (const __half*)p->residual; tp.res_ld = p->res_ld;
  tp.blend_x = (const __half*)p->blend_x; tp.blend_ld = p->blend_ld; tp.alpha = p->alpha;
  tp.out = (__half*)p->out; tp.out_ld = p->out_ld;
  if (BN == 256) return launch_tc5<256>(tp, tiles, st);
  if (BN == 128) return launch_tc5<128>(tp, tiles, st);
  return launch_tc5<64>(tp, tiles, st);
}
