// Peer-memory plumbing for the frame-sharded step: symmetric device buffers shared between the one-process-
// per-GPU ranks through CUDA IPC handles, and a flag barrier / small all-reduce kernel that works on them over NVLink.
//
// Why not NCCL for these: the frame-sharded UNet forward has ~120 tiny exchange points per step (2 per temporal
// attention, 4 per VideoResBlock).  As NCCL calls they are ~120 eager collectives that cannot overlap anything and keep
// the step out of a CUDA graph.  Here the DATA moves inside the consuming / producing kernels (peer loads in the temporal
// attention kernel, peer stores of the one-frame halo from the GroupNorm-apply kernel, see attn.cu / norm.cu); what is
// left is ordering, done by ONE single-CTA kernel per exchange point: every rank stores an epoch number into its slot of
// every peer's flag array (st.release.sys), then spins on its own array until all slots carry that epoch
// (ld.acquire.sys).  An optional payload (the [B, 32, 2] GroupNorm partial sums, <= 1024 floats) rides on the same
// kernel: written to every peer's slot before the flag, summed in rank order after the wait -- a 16-rank all-reduce in one
// launch, bit-identical on every rank.  All of it is plain stream work: capturable in the step's CUDA graph.
//
// Memory layout of the exchange area of one rank (symmetric: same layout on every rank), in 4-byte words:
//   [0, 64)                      flags[r]: last epoch rank r has announced to this rank
//   [64, 64 + 2*W*1024)          payload[parity][r][1024]
// The epoch counter is a LOCAL device word (every rank executes the same sequence of exchanges, so the counters agree);
// it is read and advanced by the kernel itself so that a replayed graph needs no new arguments.
#include <string.h>

#include "common.cuh"

namespace hi3d {

constexpr int PEER_MAX = HI3D_MAX_PEERS;
constexpr int PEER_FLAG_WORDS = 64;
constexpr int PEER_PAYLOAD_MAX = 1024;

struct PeerCtx {
  uint32_t* xchg[PEER_MAX];   // exchange area of every rank (xchg[rank] = local)
  uint32_t* epoch;            // local counter
  int rank, world;
};

HI3D_DEVINL void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
HI3D_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_exchange_kernel(const PeerCtx ctx, const float* __restrict__ payload, int n,
                                                            float* __restrict__ out) {
  __shared__ uint32_t s_epoch;
  const int tid = threadIdx.x, W = ctx.world, me = ctx.rank;
  if (tid == 0) s_epoch = *ctx.epoch + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int par = (int)(e & 1u);
  // 1. payload -> slot [par][me] of every rank (peer stores over NVLink; the local copy too)
  for (int i = tid; i < n * W; i += blockDim.x) {
    const int r = i / n, j = i - r * n;
    float* slot = reinterpret_cast<float*>(ctx.xchg[r] + PEER_FLAG_WORDS) + ((size_t)(par * W + me)) * PEER_PAYLOAD_MAX;
    slot[j] = payload[j];
  }
  __threadfence_system();          // payload (and every earlier peer store of this stream) before the flag
  __syncthreads();
  // 2. announce epoch e to every rank, 3. wait until every rank has announced e here
  if (tid < W) {
    st_release_sys(ctx.xchg[tid] + me, e);
    const uint32_t* mine = ctx.xchg[me] + tid;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if ((++spins & 1023u) == 0u) {       // deadlock watchdog (wall clock): a missing rank must not hang the GPU forever
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 90000000000ull) __trap();    // 90 s: ranks may be seconds apart (plan build, graph capture)
      }
    }
  }
  __syncthreads();
  // 4. reduce the payload in rank order (identical bits on every rank)
  if (n > 0) {
    const float* slots = reinterpret_cast<const float*>(ctx.xchg[me] + PEER_FLAG_WORDS) + (size_t)par * W * PEER_PAYLOAD_MAX;
    for (int j = tid; j < n; j += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < W; r++) s += __ldcv(slots + (size_t)r * PEER_PAYLOAD_MAX + j);
      out[j] = s;
    }
  }
  if (tid == 0) *ctx.epoch = e;
}

}  // namespace hi3d

using namespace hi3d;

extern "C" int64_t hi3d_peer_xchg_bytes(int world) {
  if (world < 1 || world > PEER_MAX) return -1;
  return (int64_t)4 * (PEER_FLAG_WORDS + 2LL * world * PEER_PAYLOAD_MAX) + 256;   // + the local epoch word (own cache line)
}

extern "C" int hi3d_symm_alloc(int64_t bytes, void** ptr, void* handle64) {
  if (bytes <= 0 || !ptr || !handle64) { set_error("hi3d_symm_alloc: bad arguments"); return -2; }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) { set_error("hi3d_symm_alloc: cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e)); return -1; }
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { set_error("hi3d_symm_alloc: %s", cudaGetErrorString(e)); cudaFree(p); return -1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int hi3d_symm_open(const void* handle64, void** ptr) {
  if (!handle64 || !ptr) { set_error("hi3d_symm_open: bad arguments"); return -2; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("hi3d_symm_open: cudaIpcOpenMemHandle: %s (peer-to-peer access between the GPUs of this box is required)",
              cudaGetErrorString(e));
    (void)cudaGetLastError();
    return -1;
  }
  *ptr = p;
  return 0;
}

extern "C" int hi3d_symm_close(void* ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { set_error("hi3d_symm_close: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

extern "C" int hi3d_symm_free(void* ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { set_error("hi3d_symm_free: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

extern "C" int hi3d_peer_exchange(void* const* xchg, int rank, int world, const float* payload, int n, float* out, void* stream) {
  if (!xchg || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || n < 0 || n > PEER_PAYLOAD_MAX ||
      (n > 0 && (!payload || !out))) {
    set_error("hi3d_peer_exchange: bad arguments (rank=%d world=%d n=%d)", rank, world, n);
    return -2;
  }
  PeerCtx ctx;
  memset(&ctx, 0, sizeof(ctx));
  for (int r = 0; r < world; r++) {
    if (!xchg[r]) { set_error("hi3d_peer_exchange: null exchange area for rank %d", r); return -2; }
    ctx.xchg[r] = (uint32_t*)xchg[r];
  }
  // the local epoch word sits behind the payload slots of the local area
  ctx.epoch = (uint32_t*)xchg[rank] + PEER_FLAG_WORDS + 2 * (size_t)world * PEER_PAYLOAD_MAX + 32;
  ctx.rank = rank; ctx.world = world;
  peer_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ctx, payload, n, out);
  return check_launch("hi3d_peer_exchange");
}
