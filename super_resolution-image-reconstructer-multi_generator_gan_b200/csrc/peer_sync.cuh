// SyncBatchNorm statistics exchange over NVLink peer memory (see peer_sync.cu).  Internal C++ interface.
#pragma once
#include <cuda_runtime.h>

#include "elementwise.cuh"

namespace srg {

struct PeerSync;
PeerSync* peer_sync_create(int world, int rank);          // allocates this rank's exchange buffer
int peer_sync_handle(PeerSync* ps, void* out64);          // cudaIpcMemHandle_t of the local buffer (64 bytes)
int peer_sync_connect(PeerSync* ps, const void* handles); // world x 64 bytes, indexed by rank
void peer_sync_destroy(PeerSync* ps);
int peer_sync_error(PeerSync* ps);                        // 1 if a wait on a peer ever timed out (synchronises)
int peer_sync_world(const PeerSync* ps);
// device-side view of the exchange buffers for kernels that fuse the exchange themselves (trunk_fused.cu)
struct PeerDeviceView {
  double* peers[8];
  int world, rank;
  unsigned long long* seq;
  int* err;
};
int peer_sync_device_view(PeerSync* ps, PeerDeviceView* out);
// partials [rows][128] -> local column sums -> exchange with all peers -> rank-ordered global sums -> finalize
// (f.count must already be the GLOBAL element count); one launch, graph capturable.
int launch_peer_finalize(PeerSync* ps, const float* partials, int rows, const ReduceFinalize& f, cudaStream_t st);

}  // namespace srg
