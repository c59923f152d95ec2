// peer_agg.cu -- the aggregate exchange step of the sharded path, fused with the local fold.
//
// sum / min / max / avg over a row-range sharded table need exactly one
// exchange step.  The NCCL form of it is three launches and two collectives (export ->
// all-reduce {sum,count} -> all-reduce {max,~min}), ~40 us on 8 GPUs for 24 bytes of
// payload -- a tenth of the 0.44 ms a 500 M-row shard's chain takes.  Here the kernel that
// folds this rank's shard partials also does the exchange, over NVLink peer memory:
//
//   * every rank owns a mailbox (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by
//     all peers): two parity banks of `world` 32-byte records;
//   * lane p of one warp (peer_exchange_warp, adb_common.cuh) stores this rank's folded partial into record [my_rank] of peer p's
//     mailbox (payload, then the epoch with st.release.sys), then spins with
//     ld.acquire.sys on record [p] of its OWN mailbox until peer p's epoch arrives, and the
//     warp folds the `world` records in rank order -- so every rank computes the same
//     result from the same operands in the same order.
//
// Two banks suffice: a peer can only start epoch e+2 after it received this rank's e+1
// record, which is sent after this rank has finished reading bank e.  All ranks must call
// in the same order (like any collective).  A rank whose peer never shows up gives up after
// kPeerTimeoutNs and reports count = -1 instead of hanging the GPU.
//
// Replaces nothing in the reference (it has no multi-device path); the values combined are
// those of sum / average / min / max, /root/reference/src/query.c:306-437.
#include "adb_common.cuh"

namespace adb {

// The stand-alone form: fold + exchange in a one-warp kernel (peer_exchange_warp,
// adb_common.cuh).  The fused chain kernel runs the same warp routine as the epilogue of its
// grid-wide aggregate fold, so a rank whose step ends with that kernel needs no extra launch.
__global__ void __launch_bounds__(kWarp)
agg_combine_allreduce_kernel(PeerExchange px) {
    peer_exchange_warp(px, (int)threadIdx.x);
}

int launch_agg_combine_allreduce(const PeerExchange &px, cudaStream_t s) {
    agg_combine_allreduce_kernel<<<1, kWarp, 0, s>>>(px);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_peer_agg() {
    preload_one(reinterpret_cast<const void *>(&agg_combine_allreduce_kernel));
}

}  // namespace adb
