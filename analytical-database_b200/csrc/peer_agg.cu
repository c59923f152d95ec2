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
//   * lane p of one warp stores this rank's folded partial into record [my_rank] of peer p's
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

constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;      // 2 s

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kWarp)
agg_combine_allreduce_kernel(const adb_agg *__restrict__ parts, int32_t k, PeerBoxes boxes,
                             int32_t rank, int32_t world, uint32_t epoch,
                             adb_agg *__restrict__ out) {
    const int lane = threadIdx.x;
    // 1. fold this rank's shard partials
    AggAcc g{0, INT32_MAX, INT32_MIN};
    int64_t cnt = 0;
    for (int i = lane; i < k; i += kWarp) {
        const adb_agg p = parts[i];
        g.sum += p.sum;
        cnt += p.count;
        g.mn = min(g.mn, p.min);
        g.mx = max(g.mx, p.max);
    }
    g.sum = warp_sum_i64(g.sum);
    cnt = warp_sum_i64(cnt);
    g.mn = warp_min_i32(g.mn);
    g.mx = warp_max_i32(g.mx);
    const uint32_t bank = epoch & 1u;
    // 2. push it into every rank's mailbox (own included: plain local stores)
    if (lane < world) {
        PeerRecord *dst = boxes.box[lane] + bank * kMaxPeers + rank;
        volatile PeerRecord *v = dst;
        v->sum = g.sum;
        v->count = cnt;
        v->min = g.mn;
        v->max = g.mx;
        st_release_sys(&dst->epoch, epoch);
    }
    // 3. wait for every rank's record in my own mailbox, fold in rank order
    AggAcc f{0, INT32_MAX, INT32_MIN};
    int64_t fc = 0;
    bool ok = true;
    if (lane < world) {
        PeerRecord *src = boxes.box[rank] + bank * kMaxPeers + lane;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&src->epoch) != epoch) {
            if (global_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
            __nanosleep(64);
        }
        const volatile PeerRecord *v = src;
        f.sum = v->sum;
        fc = v->count;
        f.mn = v->min;
        f.mx = v->max;
    }
    ok = __all_sync(kFull, ok);
    f.sum = warp_sum_i64(f.sum);
    fc = warp_sum_i64(fc);
    f.mn = warp_min_i32(f.mn);
    f.mx = warp_max_i32(f.mx);
    if (lane == 0) *out = ok ? adb_agg{f.sum, fc, f.mn, f.mx} : adb_agg{0, -1, INT32_MAX, INT32_MIN};
}

int launch_agg_combine_allreduce(const adb_agg *parts, int32_t k, const PeerBoxes &boxes, int32_t rank,
                                 int32_t world, uint32_t epoch, adb_agg *out, cudaStream_t s) {
    agg_combine_allreduce_kernel<<<1, kWarp, 0, s>>>(parts, k, boxes, rank, world, epoch, out);
    return 1;
}

}  // namespace adb
