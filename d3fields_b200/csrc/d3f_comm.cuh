// d3f_comm.cuh — peer-memory communicator for the GPUs of one box (include/d3f.h, "Multi-GPU").
//
// One cudaMalloc'ed segment per rank, mapped into every peer with CUDA IPC.  Layout of a segment:
//
//   [0, 4096)            CommHeader: epoch flags written by the peers, CTA counter, error word, broadcast flags
//   gather buffers       2 x { dist[capacity] f32 , valid[capacity] u8 }   (double-buffered by epoch parity)
//   staging              staging_bytes for d3f_comm_broadcast
//
// All synchronisation is device-side and stream-ordered: the field kernel itself publishes / awaits the epoch flags
// (gather_epilogue in d3f_common.cuh); the broadcast uses the two tiny kernels below around copy-engine P2P copies.
#pragma once
#include <cuda_runtime.h>
#include "d3f_common.cuh"

namespace d3f {

struct CommHeader {
    uint32_t gather_flag[D3F_MAX_PEERS];     // [r] = last gather epoch rank r completed (written by rank r)
    uint32_t bcast_taken[D3F_MAX_PEERS];     // [r] = last broadcast chunk rank r is done with (written by rank r to everyone)
    uint32_t bcast_ready;                    // last broadcast chunk delivered into this segment's staging (written by its root)
    uint32_t counter;                        // finished CTAs of the running gather launch (local)
    uint32_t error;                          // 1 after a timed-out wait (local)
};
static_assert(sizeof(CommHeader) <= 4096, "header must fit the first page of the segment");

// Wait until each of the `count` flags has reached `value`.
__global__ void comm_wait_kernel(const uint32_t* flags, int count, uint32_t value, uint32_t* error) {
    if (threadIdx.x < count && !wait_flag(flags + threadIdx.x, value)) *error = 1u;
}

// Publish `value` to up to D3F_MAX_PEERS remote words (after everything earlier on the stream, copies included).
struct FlagTargets { uint32_t* p[D3F_MAX_PEERS]; int32_t count; };
__global__ void comm_signal_kernel(const FlagTargets t, uint32_t value) {
    if (threadIdx.x < t.count) {
        __threadfence_system();
        st_release_sys(t.p[threadIdx.x], value);
    }
}

}  // namespace d3f

struct D3FComm {
    int rank = 0, world = 1, dev = 0;
    int64_t capacity = 0;                    // points
    size_t staging_bytes = 0, seg_bytes = 0;
    size_t off_dist[2] = {0, 0}, off_valid[2] = {0, 0}, off_staging = 0;
    char* seg[D3F_MAX_PEERS] = {};           // seg[rank] = own segment; others IPC-mapped
    bool connected = false;
    uint32_t epoch = 0;                      // gather launches issued
    uint32_t bcast_seq = 0;                  // broadcast chunks issued
    d3f::CommHeader* hdr(int r) const { return reinterpret_cast<d3f::CommHeader*>(seg[r]); }
};
