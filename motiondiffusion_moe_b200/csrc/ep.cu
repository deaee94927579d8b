// Expert-parallel MoE over NVLink peer memory (BASELINE.json configs[3]; the reference has no expert
// parallelism: its experts are a local nn.ModuleList looped in Python, models/switch_moe.py:19-25,97-109).
//
// Tokens stay sharded by sequence (data parallel); expert e of every branch lives on rank e / (E/R).
// There is no separate all-to-all: the dispatch kernel IS the token permute of the single-GPU path, it
// simply writes each LayerNorm'ed row straight into the expert-sorted buffer of the owning rank through
// a peer-mapped pointer (NVLink stores), and the combine kernel gathers the four expert rows of a token
// from the owners' output buffers (NVLink loads) while it applies the FiLM of the following
// StylizationBlock.  Row placement needs only the [R x groups] table of per-rank group counts, which
// every rank pushes to every peer (R x 64 bytes).  Cross-rank ordering uses a flag barrier in peer
// memory (st.release.sys / ld.acquire.sys, bounded spin: a lost peer sets an error word, it never hangs).
//
//   counts   (local)  per-group totals of this rank -> every peer's table; per-block bases; counters
//   barrier
//   scan     (local)  segment offsets on every owner, this rank's destination bases, local tile tables
//   dispatch (NVLink) rows + row scales -> owner buffers; perm = owner * cap + row
//   barrier
//   grouped GEMMs on the owner (mdm_gemm_bf16 with the local tile tables)
//   barrier
//   combine  (NVLink) gather + LN + FiLM + SiLU
#include "common.cuh"
#include "rowmath.cuh"

namespace {

constexpr int TOK_PER_BLK = 128;
constexpr int MAX_G = 32;

struct Topo {
  int NB, E, R, me, EPR, GT;
  __host__ __device__ int owner(int g) const { return (g % E) / EPR; }
  __host__ __device__ int local_group(int g) const { return (g / E) * EPR + (g % E) % EPR; }
};

__global__ void __launch_bounds__(32)
ep_counts_kernel(const int* __restrict__ blk_hist, const float* __restrict__ blk_imp, int nblk, Topo tp,
                 MdmEpPeers peers, int* __restrict__ blk_base, float* __restrict__ usage,
                 float* __restrict__ importance) {
  const int g = threadIdx.x;
  if (g >= tp.GT) return;
  int total = 0, top1 = 0;
  float imp = 0.f;
  for (int b = 0; b < nblk; ++b) {
    blk_base[(long)b * tp.GT + g] = total;
    total += blk_hist[((long)b * 2) * tp.GT + g];
    top1 += blk_hist[((long)b * 2 + 1) * tp.GT + g];
    imp += blk_imp[(long)b * tp.GT + g];
  }
  if (usage) usage[g] += (float)top1;
  if (importance) importance[g] += imp;
  for (int p = 0; p < tp.R; ++p) peers.cnt[p][tp.me * tp.GT + g] = total;   // 4-byte peer stores
}

__global__ void __launch_bounds__(32)
ep_scan_kernel(const int* __restrict__ cnt, Topo tp, int F, int D, int cap, int* __restrict__ dest_base,
               MTile* __restrict__ tiles_up, MTile* __restrict__ tiles_down, int* __restrict__ num_tiles,
               int* __restrict__ overflow) {
  const int g = threadIdx.x;
  const bool live = g < tp.GT;
  int rows = 0, before = 0;
  if (live) {
    for (int s = 0; s < tp.R; ++s) {
      const int c = cnt[s * tp.GT + g];
      if (s < tp.me) before += c;
      rows += c;
    }
  }
  const int padded = ((rows + 127) / 128) * 128;
  const int o = live ? tp.owner(g) : -1, lg = live ? tp.local_group(g) : 0;
  int off = 0, mine_total = 0;
  for (int j = 0; j < tp.GT; ++j) {
    const int pj = __shfl_sync(0xffffffffu, padded, j);
    const int oj = tp.owner(j), lj = tp.local_group(j);
    if (live && oj == o && lj < lg) off += pj;
    if (oj == tp.me) mine_total += pj;
  }
  if (live) {
    dest_base[g] = off + before;
    if (off + padded > cap) atomicExch(overflow, 1);
    if (o == tp.me) {
      const int ntile = padded / 128, tile0 = off / 128;
      for (int i = 0; i < ntile; ++i) {
        MTile u;
        u.a_row0 = off + i * 128; u.c_row0 = u.a_row0; u.w_row0 = lg * F; u.rows_valid = min(128, rows - i * 128);
        MTile d = u;
        d.w_row0 = lg * D;
        tiles_up[tile0 + i] = u;
        tiles_down[tile0 + i] = d;
      }
    }
  }
  if (g == 0) *num_tiles = mine_total / 128;
}

template <int VPT, typename TO>
__global__ void __launch_bounds__(256)
ep_dispatch_kernel(const float* __restrict__ x, long N, int D, Topo tp, int cap, const float* __restrict__ ln_w,
                   const float* __restrict__ ln_b, const int* __restrict__ idx, const float* __restrict__ vals,
                   const float* __restrict__ stats, const int* __restrict__ blk_base,
                   const int* __restrict__ dest_base, MdmEpPeers peers, int* __restrict__ perm) {
  __shared__ int seg_cnt[16][MAX_G];
  __shared__ int pos_s[TOK_PER_BLK * 4];
  const int NB = tp.NB, NBK = NB * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = TOK_PER_BLK * NBK;
  const int nseg = npairs / 32;
  for (int i = threadIdx.x; i < 16 * MAX_G; i += 256) (&seg_cnt[0][0])[i] = 0;
  __syncthreads();
  const long tok_blk0 = (long)blockIdx.x * TOK_PER_BLK;
  int my_g[2], my_rank[2];
  for (int r = 0; r < 2; ++r) {
    const int seg = warp * 2 + r;
    my_g[r] = -1; my_rank[r] = 0;
    if (seg < nseg) {
      const int p = seg * 32 + lane;
      const long tok = tok_blk0 + p / NBK;
      const int slot = p % NBK;
      int g = -1;
      if (tok < N) g = (slot >> 1) * tp.E + idx[tok * NBK + slot];
      const unsigned peers_m = __match_any_sync(0xffffffffu, g);
      my_g[r] = g;
      my_rank[r] = __popc(peers_m & ((1u << lane) - 1u));
      if (g >= 0 && my_rank[r] == 0) seg_cnt[seg][g] = __popc(peers_m);
    }
  }
  __syncthreads();
  for (int r = 0; r < 2; ++r) {
    const int seg = warp * 2 + r;
    if (seg < nseg && my_g[r] >= 0) {
      const int g = my_g[r];
      int base = 0;
      for (int s = 0; s < seg; ++s) base += seg_cnt[s][g];
      const int p = seg * 32 + lane;
      const long tok = tok_blk0 + p / NBK;
      const int slot = p % NBK;
      const int o = tp.owner(g);
      const int pos = dest_base[g] + blk_base[(long)blockIdx.x * tp.GT + g] + base + my_rank[r];
      pos_s[p] = o * cap + pos;
      perm[tok * NBK + slot] = o * cap + pos;
      if (pos < cap) peers.rowscale[o][pos] = vals[tok * NBK + slot] / (float)NB;
    }
  }
  __syncthreads();
  for (int it = 0; it < 16; ++it) {
    const int tl = warp * 16 + it;
    const long tok = tok_blk0 + tl;
    if (tok >= N) break;
    float v[VPT];
    load_row<VPT, float>(x + tok * D, lane, v);
    const float mean = stats[tok * 2], rstd = stats[tok * 2 + 1];
    for (int br = 0; br < NB; ++br) {
      float hrow[VPT];
#pragma unroll
      for (int i = 0; i < VPT; ++i) hrow[i] = v[i];
      affine_row<VPT>(hrow, mean, rstd, ln_w + br * D, ln_b + br * D, lane);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int code = pos_s[tl * NBK + br * 2 + k];
        const int o = code / cap, pos = code - o * cap;
        if (pos < cap) store_row<VPT, TO>(reinterpret_cast<TO*>(peers.xp[o]) + (long)pos * D, lane, hrow);
      }
    }
  }
}

template <int VPT, typename TI>
__global__ void __launch_bounds__(256)
ep_combine_film_kernel(MdmEpPeers peers, const int* __restrict__ perm, long N, int D, int NBK, int cap,
                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                       const float* __restrict__ film, int rows_per_seq, TI* __restrict__ out) {
  const long tok = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= N) return;
  const int lane = threadIdx.x & 31;
  float acc[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[i] = 0.f;
  for (int br = 0; br < NBK / 2; ++br) {
    float a[VPT], b[VPT];
    const int c0 = perm[tok * NBK + br * 2], c1 = perm[tok * NBK + br * 2 + 1];
    const int o0 = c0 / cap, o1 = c1 / cap;
    load_row<VPT, TI>(reinterpret_cast<const TI*>(peers.yp[o0]) + (long)(c0 - o0 * cap) * D, lane, a);
    load_row<VPT, TI>(reinterpret_cast<const TI*>(peers.yp[o1]) + (long)(c1 - o1 * cap) * D, lane, b);
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc[i] += a[i] + b[i];
  }
  layernorm_row<VPT>(acc, ln_w, ln_b, lane, D);
  film_row<VPT>(acc, film + (tok / rows_per_seq) * 2 * D, lane, D);
#pragma unroll
  for (int i = 0; i < VPT; ++i) acc[i] = silu_out<TI>(acc[i]);
  store_row<VPT, TI>(out + tok * D, lane, acc);
}

// Flag barrier across the R ranks of one node.  Thread p publishes `epoch` into slot [me] of rank p's
// flag array (release, system scope) and then waits until rank p has published >= epoch into this
// rank's slot [p].  Epochs only grow (a device-side counter per rank), so no reset is needed.  The spin is bounded (~2 s): on expiry the
// error word is set and the kernel traps (fatal, loud), so a dead peer can neither hang the GPU nor corrupt a sample.
__global__ void __launch_bounds__(32)
ep_barrier_kernel(MdmEpPeers peers, int R, int me, unsigned* __restrict__ epoch_ctr, int* __restrict__ err) {
  const int p = threadIdx.x;
  // the epoch lives in device memory and advances by one per barrier, so that a captured CUDA graph
  // (fixed kernel arguments) can be replayed: every rank executes the same sequence of barriers
  unsigned epoch = 0;
  if (p == 0) epoch = ++(*epoch_ctr);
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  __threadfence_system();
  if (p < R) {
    unsigned* remote = peers.flags[p] + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned* local = peers.flags[me] + p;
    unsigned v = 0;
    long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
      if ((int)(v - epoch) >= 0) break;
      __nanosleep(200);
    } while (++spins < 40000000L);   // ~10 s
    if ((int)(v - epoch) < 0) {
      // a peer did not arrive: its buffers are not ready, so nothing after this barrier may run on them.  Record the
      // reason, then abort the grid: the stream fails with a launch error at the next synchronisation instead of
      // producing silently wrong samples (the error is sticky for the context, i.e. fatal for this process).
      atomicExch(err, 1);
      __threadfence_system();
      __trap();
    }
  }
  __syncwarp();
  __threadfence_system();
}

Topo make_topo(int NB, int E, int R, int me) {
  Topo t;
  t.NB = NB; t.E = E; t.R = R; t.me = me; t.EPR = E / R; t.GT = NB * E;
  return t;
}
bool topo_ok(int NB, int E, int K, int R, int me) {
  return K == 2 && NB >= 1 && NB <= 2 && R >= 1 && R <= MDM_EP_MAX_RANKS && me >= 0 && me < R && E % R == 0 &&
         NB * E <= MAX_G;
}

}  // namespace

#define VPT_SWITCH(D, ...)               \
  switch (D) {                           \
    case 128: { constexpr int V = 4; __VA_ARGS__; break; }  \
    case 256: { constexpr int V = 8; __VA_ARGS__; break; }  \
    case 512: { constexpr int V = 16; __VA_ARGS__; break; } \
    case 1024: { constexpr int V = 32; __VA_ARGS__; break; }\
    default: return MDM_ERR_UNSUPPORTED; \
  }

extern "C" MDM_API int mdm_ep_counts(const int* blk_hist, const float* blk_imp, long N, int NB, int E, int K,
                                     int R, int me, const MdmEpPeers* peers, int* blk_base, float* usage,
                                     float* importance, void* stream) {
  if (!blk_hist || !blk_imp || !peers || !blk_base) return MDM_ERR_ARG;
  if (!topo_ok(NB, E, K, R, me)) return MDM_ERR_UNSUPPORTED;
  const int nblk = (int)((N + TOK_PER_BLK - 1) / TOK_PER_BLK);
  ep_counts_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(blk_hist, blk_imp, nblk, make_topo(NB, E, R, me),
                                                                         *peers, blk_base, usage, importance);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_ep_scan(const int* cnt, int NB, int E, int K, int R, int me, int F, int D, int cap,
                                   int* dest_base, void* tiles_up, void* tiles_down, int* num_tiles,
                                   int* overflow, void* stream) {
  if (!cnt || !dest_base || !tiles_up || !tiles_down || !num_tiles || !overflow) return MDM_ERR_ARG;
  if (!topo_ok(NB, E, K, R, me) || (cap & 127)) return MDM_ERR_UNSUPPORTED;
  ep_scan_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cnt, make_topo(NB, E, R, me), F, D, cap, dest_base, reinterpret_cast<MTile*>(tiles_up),
      reinterpret_cast<MTile*>(tiles_down), num_tiles, overflow);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_ep_dispatch(const float* x, long N, int D, int NB, int E, int K, int R, int me, int cap,
                                       const float* ln_w, const float* ln_b, const int* idx, const float* vals,
                                       const float* stats, const int* blk_base, const int* dest_base,
                                       const MdmEpPeers* peers, int dt, int* perm, void* stream) {
  if (!x || !ln_w || !ln_b || !idx || !vals || !stats || !blk_base || !dest_base || !peers || !perm) return MDM_ERR_ARG;
  if (!topo_ok(NB, E, K, R, me)) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  const int nblk = (int)((N + TOK_PER_BLK - 1) / TOK_PER_BLK);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Topo tp = make_topo(NB, E, R, me);
  VPT_SWITCH(D, {
    if (dt == MDM_F32)
      ep_dispatch_kernel<V, float><<<nblk, 256, 0, st>>>(x, N, D, tp, cap, ln_w, ln_b, idx, vals, stats, blk_base,
                                                          dest_base, *peers, perm);
    else
      ep_dispatch_kernel<V, bf16><<<nblk, 256, 0, st>>>(x, N, D, tp, cap, ln_w, ln_b, idx, vals, stats, blk_base,
                                                         dest_base, *peers, perm);
  });
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_ep_combine_film(const MdmEpPeers* peers, int dt, const int* perm, long N, int D, int NBK,
                                           int cap, const float* ln_w, const float* ln_b, const float* film,
                                           int rows_per_seq, void* out, void* stream) {
  if (!peers || !perm || !ln_w || !ln_b || !film || !out || rows_per_seq <= 0 || cap <= 0) return MDM_ERR_ARG;
  if (NBK < 2 || (NBK & 1)) return MDM_ERR_UNSUPPORTED;
  if (N == 0) return MDM_OK;
  const unsigned grid = (unsigned)((N + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VPT_SWITCH(D, {
    if (dt == MDM_F32)
      ep_combine_film_kernel<V, float><<<grid, 256, 0, st>>>(*peers, perm, N, D, NBK, cap, ln_w, ln_b, film, rows_per_seq,
                                                              reinterpret_cast<float*>(out));
    else
      ep_combine_film_kernel<V, bf16><<<grid, 256, 0, st>>>(*peers, perm, N, D, NBK, cap, ln_w, ln_b, film, rows_per_seq,
                                                             reinterpret_cast<bf16*>(out));
  });
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

extern "C" MDM_API int mdm_ep_barrier(const MdmEpPeers* peers, int R, int me, unsigned* epoch_ctr, int* err,
                                      void* stream) {
  if (!peers || !err || !epoch_ctr || R < 1 || R > MDM_EP_MAX_RANKS || me < 0 || me >= R) return MDM_ERR_ARG;
  ep_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*peers, R, me, epoch_ctr, err);
  return cudaGetLastError() == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}

// ---- CUDA IPC plumbing for the peer-mapped buffers (one process per GPU) -------------------------
extern "C" MDM_API int mdm_ipc_get_handle(const void* ptr, void* handle64, long* offset) {
  if (!ptr || !handle64 || !offset) return MDM_ERR_ARG;
  // the handle names the whole allocation: report where `ptr` sits inside it
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return MDM_ERR_CUDA;
  CUdeviceptr base = 0;
  size_t size = 0;
  if (reinterpret_cast<RangeFn>(fp)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return MDM_ERR_CUDA;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)) != cudaSuccess) return MDM_ERR_CUDA;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (long)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return MDM_OK;
}
extern "C" MDM_API int mdm_ipc_open_handle(const void* handle64, void** base) {
  if (!handle64 || !base) return MDM_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
extern "C" MDM_API int mdm_ipc_close_handle(void* base) {
  return cudaIpcCloseMemHandle(base) == cudaSuccess ? MDM_OK : MDM_ERR_CUDA;
}
