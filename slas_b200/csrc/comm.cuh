// comm.cuh -- internal face of comm.cu
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace slas {
int comm_nranks();
// In-place sum-allreduce of `count` f64 values at dev_inout over all ranks, on stream st.
// A no-op when no communicator (or a 1-rank one) is active.
int comm_allreduce_f64(double *dev_inout, size_t count, cudaStream_t st);
// Peer-memory exchange (p2p.cuh): true + a descriptor with a fresh epoch when the NVLink mailboxes are mapped.
bool comm_next_peer_exchange(PeerExchange *out);
// The launch that was to use `epoch` failed before reaching the GPU: give the epoch back so that this rank stays
// in step with its peers.
void comm_rollback_peer_exchange(unsigned long long epoch);
}  // namespace slas
