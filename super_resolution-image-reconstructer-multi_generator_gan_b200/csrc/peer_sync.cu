// SyncBatchNorm statistics exchange over NVLink peer memory, fused with the reduction and the BatchNorm finalize.
//
// One 1024-thread block per rank: (1) fixed-order column sums of the local per-CTA partials, (2) the 128 doubles are
// written straight into EVERY peer's exchange buffer (P2P stores through NVSwitch) followed by a release-flag,
// (3) the block waits until all peers' flags for this collective arrived in its own buffer, (4) sums the `world`
// contributions in rank order (bit-identical on every rank) and (5) runs the per-channel finalize.  Compared with
// "reduce kernel -> ncclAllReduce -> finalize kernel" this is one launch and one NVLink round trip (64 such collectives
// sit on the critical path of every generator forward+backward).  The collective sequence number lives in device memory
// and is advanced by the kernel itself, so the launch is CUDA-graph capturable.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "launch.cuh"
#include "peer_sync.cuh"

namespace srg {

namespace {
constexpr int kSlots = 4;
constexpr int kSlotDoubles = 136;   // 128 sums + flag + padding (1088 bytes)
constexpr int kMaxWorld = 8;
// ~64 ns per probe: 2^26 probes is minutes, long enough for a peer held up by data loading, checkpoint I/O or a first-time
// graph capture (the round-1 bound of 2^22 gave up after a few seconds and then normalised with stale sums)
constexpr long long kPeerWaitSpins = 1ll << 26;

struct PeerParams {
  double* peers[kMaxWorld];   // exchange buffers of all ranks (peers[rank] is the local one)
  int world, rank;
  unsigned long long* seq;    // device counter of collectives issued on this communicator
  int* err;                   // set to 1 if a wait timed out
  int ll;                     // 1: data and sequence number travel in the same 8-byte words (no fence, no separate flag)
};

// "LL" exchange region behind the flag-protocol region of every rank's buffer: per (slot, source rank) 128 doubles as 256
// words {u32 half of the double, u32 sequence number}.  An aligned 8-byte store is single-copy atomic, so a word whose
// sequence half matches carries valid data: the sender needs no system-scope fence and no separate flag store, the receiver no
// second round trip (the scheme of NCCL's low-latency protocol).
constexpr int kLLWords = 256;
__device__ __forceinline__ unsigned long long* ll_region(double* base, int world) {
  return reinterpret_cast<unsigned long long*>(base + size_t(kSlots) * world * kSlotDoubles);
}
__device__ __forceinline__ void st_volatile_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void finalize_channels_peer(const ReduceFinalize& f, int c, double s1, double s2, float pscale) {
  if (f.mode == RF_BN_FWD) {
    const double mean = s1 / f.count;
    double var = s2 / f.count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = float(1.0 / sqrt(var + double(f.eps)));
    const float sc = f.gamma[c] * inv;
    f.out0[c] = sc;
    f.out1[c] = f.beta[c] - float(mean) * sc;
    f.out2[c] = float(mean);
    f.out3[c] = inv;
    if (f.running_mean != nullptr) {
      const double unbiased = f.count > 1.0 ? var * (f.count / (f.count - 1.0)) : var;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * float(mean);
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * float(unbiased);
    }
  } else {
    const double mean = f.save_mean[c], inv = f.save_inv[c];
    const double dg = inv * (s2 - mean * s1);
    const double db = s1;
    if (f.dgamma) f.dgamma[c] = float(dg) * pscale;   // 1/world: see bn_bwd_finalize_kernel
    if (f.dbeta) f.dbeta[c] = float(db) * pscale;
    const double sc = double(f.gamma[c]) * inv;
    f.out0[c] = float(sc);
    f.out1[c] = float(-sc * inv * dg / f.count);
    f.out2[c] = float(-sc * db / f.count + sc * inv * mean * dg / f.count);
  }
}

__global__ void __launch_bounds__(1024) peer_finalize_kernel(const float* __restrict__ partials, int rows, const ReduceFinalize f,
                                                             const PeerParams pp) {
  __shared__ double red[8][128];
  __shared__ unsigned long long seq_s;
  __shared__ int timed_out;
  // NOTE: pdl_trigger() only AFTER the cross-GPU wait below.  Triggering early would let the dependent kernel's blocks
  // fill every SM while this block spins on a peer; the peer's matching kernel of ANOTHER graph branch could then find
  // no free SM on its GPU and the two GPUs would wait on each other forever.
  pdl_wait();
  const int col = threadIdx.x & 127, rl = threadIdx.x >> 7;
  red[rl][col] = partials_lane_sum(partials, rows, col, rl);
  if (threadIdx.x == 0) { seq_s = ++(*pp.seq); timed_out = 0; }
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int slot = int(seq % kSlots);
  if (pp.ll) {
    // (2)-(4) in one round: every column thread sends its sum to every rank as two self-validating words and collects the
    // ranks' words for its column from the local buffer, in rank order
    if (rl == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][col];
      const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(t));
      const unsigned long long tag = (seq & 0xffffffffull) << 32;
      const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
      const size_t mine = (size_t(slot) * pp.world + pp.rank) * kLLWords + 2 * col;
      for (int r = 0; r < pp.world; ++r) st_volatile_v2(ll_region(pp.peers[r], pp.world) + mine, w0, w1);
      double total = 0.0;
      bool lost = false;
      const unsigned long long* in = ll_region(pp.peers[pp.rank], pp.world) + size_t(slot) * pp.world * kLLWords + 2 * col;
      for (int r = 0; r < pp.world; ++r) {
        unsigned long long a, b;
        long long spins = 0;
        for (;;) {
          ld_volatile_v2(in + size_t(r) * kLLWords, a, b);
          if ((a >> 32) == (tag >> 32) && (b >> 32) == (tag >> 32)) break;
          if (++spins > kPeerWaitSpins) { lost = true; break; }
          if (spins > 64) __nanosleep(32);
        }
        total += __longlong_as_double(static_cast<long long>((a & 0xffffffffull) | (b << 32)));
      }
      if (lost) { *pp.err = 1; total = __longlong_as_double(0x7ff8000000000000ll); }   // NaN: never silently use a missing contribution
      red[0][col] = total;     // (red[0][col] was read above by this thread only)
    }
    __syncthreads();
    pdl_trigger();
  } else {
  // (2) my 128 sums -> every rank's buffer, slot [slot][my rank]
    if (rl == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][col];
      for (int r = 0; r < pp.world; ++r) pp.peers[r][(size_t(slot) * pp.world + pp.rank) * kSlotDoubles + col] = t;
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < pp.world) {
      const int r = threadIdx.x;
      st_release_sys(reinterpret_cast<unsigned long long*>(pp.peers[r] + (size_t(slot) * pp.world + pp.rank) * kSlotDoubles + 128), seq);
      // (3) wait for rank r's contribution in my own buffer
      const unsigned long long* flag =
          reinterpret_cast<const unsigned long long*>(pp.peers[pp.rank] + (size_t(slot) * pp.world + r) * kSlotDoubles + 128);
      long long spins = 0;
      while (ld_acquire_sys(flag) != seq) {
        // a lost / stalled peer must never hang the GPU -- but it must not go unnoticed either: the statistics are
        // poisoned below (NaN losses on every rank that missed a contribution) and *err makes the host raise
        if (++spins > kPeerWaitSpins) { *pp.err = 1; timed_out = 1; break; }
        __nanosleep(64);
      }
    }
    __syncthreads();
    pdl_trigger();
    // (4) rank-ordered total, (5) finalize
    if (rl == 0) {
      double t = 0.0;
      for (int r = 0; r < pp.world; ++r) t += ld_volatile_f64(pp.peers[pp.rank] + (size_t(slot) * pp.world + r) * kSlotDoubles + col);
      if (timed_out) t = __longlong_as_double(0x7ff8000000000000ll);      // NaN: a missing contribution is never silently used
      red[0][col] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) finalize_channels_peer(f, threadIdx.x, red[0][threadIdx.x], red[0][64 + threadIdx.x], 1.f / float(pp.world));
}

}  // namespace

struct PeerSync {
  int world = 1, rank = 0;
  double* local = nullptr;
  double* peers[kMaxWorld] = {};
  bool opened[kMaxWorld] = {};
  unsigned long long* d_seq = nullptr;
  int* d_err = nullptr;
};

PeerSync* peer_sync_create(int world, int rank) {
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) { set_error("peer_sync_create: bad world/rank"); return nullptr; }
  PeerSync* ps = new PeerSync();
  ps->world = world; ps->rank = rank;
  // flag-protocol region (also used by the fused trunk kernel's in-kernel exchange) + LL region (peer_finalize_kernel)
  const size_t bytes = size_t(kSlots) * world * kSlotDoubles * sizeof(double) + size_t(kSlots) * world * kLLWords * 8;
  if (cudaMalloc(&ps->local, bytes) != cudaSuccess || cudaMalloc(&ps->d_seq, 8) != cudaSuccess ||
      cudaMalloc(&ps->d_err, 4) != cudaSuccess) {
    set_error("peer_sync_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ps;
    return nullptr;
  }
  cudaMemset(ps->local, 0, bytes);
  cudaMemset(ps->d_seq, 0, 8);
  cudaMemset(ps->d_err, 0, 4);
  ps->peers[rank] = ps->local;
  return ps;
}
int peer_sync_handle(PeerSync* ps, void* out64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ps->local);
  if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return int(e); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out64, &h, 64);
  return 0;
}
int peer_sync_connect(PeerSync* ps, const void* handles) {
  for (int r = 0; r < ps->world; ++r) {
    if (r == ps->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + size_t(r) * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)); return int(e); }
    ps->peers[r] = static_cast<double*>(ptr);
    ps->opened[r] = true;
  }
  return 0;
}
void peer_sync_destroy(PeerSync* ps) {
  if (ps == nullptr) return;
  for (int r = 0; r < ps->world; ++r)
    if (ps->opened[r]) cudaIpcCloseMemHandle(ps->peers[r]);
  cudaFree(ps->local); cudaFree(ps->d_seq); cudaFree(ps->d_err);
  delete ps;
}
int peer_sync_error(PeerSync* ps) {
  int v = 0;
  cudaMemcpy(&v, ps->d_err, 4, cudaMemcpyDeviceToHost);
  return v;
}
int peer_sync_world(const PeerSync* ps) { return ps->world; }
int peer_sync_device_view(PeerSync* ps, PeerDeviceView* out) {
  memset(out, 0, sizeof(*out));
  for (int r = 0; r < ps->world; ++r) {
    if (ps->peers[r] == nullptr) { set_error("peer_sync: rank %d not connected", r); return -2; }
    out->peers[r] = ps->peers[r];
  }
  out->world = ps->world; out->rank = ps->rank; out->seq = ps->d_seq; out->err = ps->d_err;
  return 0;
}

int launch_peer_finalize(PeerSync* ps, const float* partials, int rows, const ReduceFinalize& f, cudaStream_t st) {
  if (f.mode != RF_BN_FWD && f.mode != RF_BN_BWD) { set_error("peer_finalize: unsupported mode"); return -1; }
  PeerParams pp;
  memset(&pp, 0, sizeof(pp));
  for (int r = 0; r < ps->world; ++r) {
    if (ps->peers[r] == nullptr) { set_error("peer_finalize: rank %d not connected", r); return -2; }
    pp.peers[r] = ps->peers[r];
  }
  pp.world = ps->world; pp.rank = ps->rank; pp.seq = ps->d_seq; pp.err = ps->d_err;
  // SRG_PEER_LL=0: the round-1 protocol (data, system-scope fence, flag store, flag poll, data read) for A/B runs
  static const int ll = [] { const char* v = getenv("SRG_PEER_LL"); return (v != nullptr && v[0] == '0') ? 0 : 1; }();
  pp.ll = ll;
  cudaError_t e = launch_pdl(peer_finalize_kernel, dim3(1), dim3(1024), 0, st, partials, rows, f, pp);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("peer_finalize launch: %s", cudaGetErrorString(e)); return int(e); }
  count_launch();
  return 0;
}

}  // namespace srg
