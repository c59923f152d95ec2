// peer_exchange.cu -- the pair exchange of the sharded hash join, without a collective library.
//
// the hash join shards by hash-partitioning both (value, position) pair lists
// over the ranks -- route, all-to-all-v, local build + probe.  The NCCL form is six collectives
// and four host synchronisations per join (counts, values, positions, for both sides).  Here
// the routing kernel IS the exchange:
//
//   rx_hist + row scan    per-destination pair counts of this rank, per-tile offsets (radix.cu);
//   jx_counts_kernel      one warp: stores this rank's count row into every rank's mailbox,
//                         acquire-spins until all rows of this epoch have arrived (an all-gather
//                         that doubles as the "previous receive buffer has been consumed"
//                         barrier: nobody gets past it before everybody has enqueued it behind
//                         its own previous join), then derives where this rank's piece starts
//                         inside every destination's receive region (sum of the lower ranks'
//                         counts: pieces land ordered by source rank, in source order inside a
//                         piece, exactly the layout of an all-to-all-v) and how many pairs this
//                         rank will receive;
//   rx_scatter<REMOTE>    the stable routing scatter, writing each destination's run straight
//                         into that rank's receive region over NVLink;
//   jx_done_kernel        one warp: fence, done flag to every rank, wait for every source's flag.
//
// Receive regions are persistent (adb_peer_join_create: one cudaMalloc per rank, mapped by all
// peers through CUDA IPC): registering a fresh buffer per join would cost more than the join.
// A region that would overflow aborts the exchange on every rank alike (they all see the same
// count matrix); a peer that never arrives is given up on after 2 s.
#include "adb_common.cuh"

namespace adb {

__global__ void __launch_bounds__(kWarp)
jx_counts_kernel(PairExchange x, const uint32_t *__restrict__ my_totals, uint32_t *__restrict__ base_out,
                 int64_t *__restrict__ recv_total_out, uint32_t *__restrict__ status) {
    const int lane = threadIdx.x;
    const uint32_t bank = x.epoch & 1u;
    // my count row -> every rank's matrix
    if (lane < x.world) {
        PeerCtl *ctl = reinterpret_cast<PeerCtl *>(x.boxes->box[lane]);
        volatile uint32_t *row = ctl->jcnt[bank][x.rank];
        for (int d = 0; d < x.world; ++d) row[d] = my_totals[d];
        st_release_sys(&ctl->jcnt_epoch[bank][x.rank], x.epoch);
    }
    // wait for every source's row in my own matrix
    PeerCtl *mine = reinterpret_cast<PeerCtl *>(x.boxes->box[x.rank]);
    bool ok = true;
    if (lane < x.world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&mine->jcnt_epoch[bank][lane]) != x.epoch) {
            if (global_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
            __nanosleep(32);
        }
    }
    ok = __all_sync(kFull, ok);
    // lane d: where my piece starts in destination d, and how much d receives in total
    uint32_t before = 0;
    unsigned long long total = 0;
    if (ok && lane < x.world) {
        for (int s = 0; s < x.world; ++s) {
            const uint32_t c = *reinterpret_cast<volatile uint32_t *>(&mine->jcnt[bank][s][lane]);
            if (s < x.rank) before += c;
            total += c;
        }
    }
    const bool over = __any_sync(kFull, total > x.cap);
    if (lane < x.world) base_out[lane] = before;
    if (lane == x.rank) *recv_total_out = (int64_t)total;
    if (lane == 0) status[0] = !ok ? 2u : over ? 1u : 0u;
}

__global__ void __launch_bounds__(kWarp)
jx_done_kernel(PairExchange x, uint32_t *__restrict__ status) {
    const int lane = threadIdx.x;
    const uint32_t bank = x.epoch & 1u;
    __threadfence_system();
    if (lane < x.world) {
        PeerCtl *ctl = reinterpret_cast<PeerCtl *>(x.boxes->box[lane]);
        st_release_sys(&ctl->jdone_epoch[bank][x.rank], x.epoch);
    }
    PeerCtl *mine = reinterpret_cast<PeerCtl *>(x.boxes->box[x.rank]);
    bool ok = true;
    if (lane < x.world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&mine->jdone_epoch[bank][lane]) != x.epoch) {
            if (global_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
            __nanosleep(32);
        }
    }
    ok = __all_sync(kFull, ok);
    if (lane == 0 && !ok) status[0] = 2u;
}

int launch_jx_counts(const PairExchange &x, const uint32_t *my_totals, uint32_t *base_out,
                     int64_t *recv_total_out, uint32_t *status, cudaStream_t s) {
    jx_counts_kernel<<<1, kWarp, 0, s>>>(x, my_totals, base_out, recv_total_out, status);
    return 1;
}
int launch_jx_done(const PairExchange &x, uint32_t *status, cudaStream_t s) {
    jx_done_kernel<<<1, kWarp, 0, s>>>(x, status);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_peer_exchange() {
    preload_one(reinterpret_cast<const void *>(&jx_counts_kernel));
    preload_one(reinterpret_cast<const void *>(&jx_done_kernel));
}

}  // namespace adb
