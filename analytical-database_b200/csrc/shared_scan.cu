// shared_scan.cu -- batch_queries / batch_execute: Q range selects in one pass over a column.
//
// Replaces shared_select + select_task (/root/reference/src/query.c:450-583), whose inner
// loop tests every row against every query (N x Q comparisons on 3 pthreads).  On a GPU
// that loop is ALU-bound long before it is HBM-bound (100 M rows x 100 queries = 2e10
// compares), so the batch is evaluated differently:
//
//   host       the <= 2Q distinct bounds cut the value domain into elementary intervals;
//              every interval knows the (few) queries that cover it (a CSR list, query
//              ids ascending);
//   classify   one streaming pass: a 64 K-bucket bitmap in shared memory dismisses rows no
//              query wants with one bit test; candidate rows are compacted per warp and
//              resolved densely (exact interval, per-warp-chunk per-interval counters) into
//              the chunk's hit list;
//   offsets    one CTA per query scans its row of the [query][chunk] count matrix;
//   emit       each warp turns its chunk's hit list into the queries' position lists.  Every
//              list comes out ascending, as the reference's memcpy-concatenated thread
//              slices do (query.c:563-574).
//
// The hit lists come in two forms.  While no value is wanted by more than four queries (the
// usual batch) they hold one {query, row} entry per (row, query) pair and the emit pass sorts
// them by query, tile by tile, in shared memory (ss_emit_sorted_kernel).  Heavily overlapping
// batches list one {interval, row} entry per hit row and the emit pass expands each row's
// cover list through a small pair queue (ss_emit_kernel).
//
// has_low / has_high are ignored and all queries share one column, exactly as
// query.c:474 and server.c:376 do.
#include "adb_common.cuh"

namespace adb {

constexpr int SS_THREADS = 256;
constexpr int SS_WARPS = SS_THREADS / kWarp;
constexpr int SS_WTILE = 512;                     // rows per warp-tile (same layout as select)
constexpr int SS_QMAX = ADB_MAX_BATCH;            // 150
constexpr int SS_BMAX = 2 * SS_QMAX;              // <= 300 distinct bounds
constexpr int SS_LUT = 1024;                      // value -> first candidate interval
constexpr int SS_QUEUE = 256;                     // (query, row) pairs parked per warp (power of two)

// Interval id of v = number of bounds <= v, in [0, m].  A 1024-entry table over
// [bounds[0], bounds[m-1]) gives the id at the lower edge of v's bucket; a short forward walk
// (or a binary search inside the bucket when many bounds crowd into it) finishes the job --
// one subtraction, one shift and two or three shared-memory reads for typical batches,
// against log2(2Q) dependent probes for a plain binary search.  The table is built once on
// the host (engine.cu) and copied into shared memory by every CTA.
struct SsLookup {
    const int32_t *bounds;
    const uint16_t *lut;                          // SS_LUT + 1 entries
    uint32_t m, shift;
    int32_t lo, hi;                               // bounds[0], bounds[m-1]
};

__device__ __forceinline__ uint32_t interval_of(const SsLookup &L, int32_t v) {
    if (L.m == 0 || v < L.lo) return 0u;
    if (v >= L.hi) return L.m;
    const uint32_t k = ((uint32_t)v - (uint32_t)L.lo) >> L.shift;
    uint32_t id = L.lut[k];
    if (id & 0x8000u) return 0u;                  // no query reaches into this bucket: "no hit"
    const uint32_t id_end = L.lut[k + 1] & 0x7FFFu;   // every bound of this bucket is < id_end
    if (id_end - id <= 4) {
        while (id < id_end && L.bounds[id] <= v) ++id;
        return id;
    }
    uint32_t lo = id, hi = id_end;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (L.bounds[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Classify + count.  The common case of a selective batch is a row no query wants: it is
// dismissed branch-free with one bit test in a 64 K-bucket bitmap over the value domain
// (8 KB of shared memory; a bucket is set when some query reaches into it, so what survives
// is the hits plus the rows in the two edge buckets of every query -- r01l's 1024-bucket
// table let twice as many rows through as really hit).  Survivors are compacted, in row
// order, into a per-warp work list that is only resolved when it cannot take another 128
// rows, so the expensive part always runs on full warps: exact interval, per-query
// counters, and -- if at least one query covers the interval -- one packed
// {interval, row} entry appended to the chunk's hit list in global memory.  The emit pass
// reads that list instead of the column.
constexpr int SS_WORK = SS_WTILE;                 // work-list entries per warp
constexpr int SS_BITWORDS = kSsBits / 32;         // 2048
static_assert(SS_LUT == (int)kSsLut, "host and device agree on the table size");

struct SsShared {                                  // carved out of dynamic shared memory
    uint32_t bits[SS_BITWORDS + 4];                // [SS_BITWORDS] stays 0: where values past the last bound land
    uint2 work[SS_WARPS][SS_WORK];
    uint32_t cnt[SS_WARPS][SS_BMAX + 4];           // hits per INTERVAL id (0 .. m), per warp
    int32_t bounds[SS_BMAX];
    uint16_t lut[SS_LUT + 2];
    uint32_t pack[SS_BMAX + 2];                    // per interval id: pair lists -- its covering queries by colour, one
                                                   // byte each, 0xFF = none; interval lists -- 0 = covered, ~0 = not
};

// resolve work[0 .. wcount): all lanes busy except in the last batch.
// PAIRS: one {query, row} entry per (row, covering query) pair, written colour by colour: the
// host coloured the queries so that overlapping ones differ, a query therefore lives in one
// colour and its entries stay in row order -- all the emit pass needs -- and a batch costs
// one ballot-compaction per colour instead of a scan plus a loop over each row's cover list.
// Otherwise one {interval, row} entry per hit row.
template <bool PAIRS>
__device__ __forceinline__ uint32_t ss_resolve(const SsLookup &L, const uint32_t *s_pack, uint32_t depth,
                                               const uint2 *work, uint32_t wcount, uint32_t *my_cnt,
                                               uint32_t *__restrict__ my_hits, uint32_t nhits,
                                               uint32_t lane, uint32_t lt) {
    for (uint32_t base = 0; base < wcount; base += kWarp) {
        const bool live = base + lane < wcount;
        uint32_t id = 0, rel = 0, p = 0xFFFFFFFFu;
        if (live) {
            const uint2 w = work[base + lane];
            rel = w.y;
            id = interval_of(L, (int32_t)w.x);
            p = s_pack[id];
        }
        // hits are counted per interval here; a query covers a contiguous run of intervals, so
        // its count is a difference of two prefix sums, taken once per chunk
        if (p != 0xFFFFFFFFu) atomicAdd(&my_cnt[id], 1u);
        if (!PAIRS) {
            const uint32_t m = __ballot_sync(kFull, p != 0xFFFFFFFFu);
            if (p != 0xFFFFFFFFu) my_hits[nhits + __popc(m & lt)] = (id << 23) | rel;   // rel < 2^23, id <= 300
            nhits += __popc(m);
        } else {
            for (uint32_t c = 0; c < depth; ++c) {
                const uint32_t q = (p >> (8 * c)) & 0xFFu;
                const uint32_t m = __ballot_sync(kFull, q != 0xFFu);
                if (q != 0xFFu) my_hits[nhits + __popc(m & lt)] = (q << 23) | rel;
                nhits += __popc(m);
            }
        }
    }
    return nhits;
}

template <bool PAIRS>
__global__ void __launch_bounds__(SS_THREADS)
ss_classify_kernel(const int32_t *__restrict__ val, uint32_t n, SharedScanPlan plan,
                   uint32_t chunk_rows, uint32_t num_chunks, uint32_t *__restrict__ hitlist,
                   uint32_t *__restrict__ chunk_hits, uint32_t *__restrict__ counts /* [q][num_chunks] */) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    SsShared &S = *reinterpret_cast<SsShared *>(ss_smem);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < SS_BITWORDS / 4; i += SS_THREADS)
        reinterpret_cast<uint4 *>(S.bits)[i] = reinterpret_cast<const uint4 *>(plan.bits)[i];
    for (uint32_t i = threadIdx.x; i < SS_LUT + 1; i += SS_THREADS) S.lut[i] = plan.lut[i];
    for (uint32_t i = threadIdx.x; i < plan.m; i += SS_THREADS) S.bounds[i] = plan.bounds[i];
    for (uint32_t i = threadIdx.x; i <= plan.m; i += SS_THREADS)
        S.pack[i] = PAIRS ? plan.cov4[i] : (plan.cov_off[i + 1] > plan.cov_off[i] ? 0u : 0xFFFFFFFFu);
    for (uint32_t i = threadIdx.x; i < SS_WARPS * (SS_BMAX + 4); i += SS_THREADS) (&S.cnt[0][0])[i] = 0;
    if (threadIdx.x < 4) S.bits[SS_BITWORDS + threadIdx.x] = 0;
    __syncthreads();
    const SsLookup L{S.bounds, S.lut, plan.m, plan.lut_shift, plan.lo, (int32_t)((uint32_t)plan.lo + plan.span)};
    const uint32_t chunk = blockIdx.x * SS_WARPS + warp;
    if (chunk >= num_chunks) return;
    const uint32_t row_begin = chunk * chunk_rows;
    const uint32_t tiles = chunk_rows / SS_WTILE;
    const bool aligned = (reinterpret_cast<uintptr_t>(val) & 15u) == 0;
    uint32_t *my_cnt = S.cnt[warp];
    uint2 *work = S.work[warp];
    const uint32_t *bits = S.bits;
    uint32_t *__restrict__ my_hits = hitlist + (size_t)chunk * chunk_rows * (PAIRS ? plan.pair_depth : 1u);
    const uint32_t ulo = (uint32_t)plan.lo, bsh = plan.bit_shift;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t nhits = 0, wcount = 0;                             // uniform across the warp

    for (uint32_t t = 0; t < tiles; ++t) {
        const uint32_t row0 = row_begin + t * SS_WTILE;
        if (row0 >= n) break;
        int32_t v[16];
        uint32_t okmask = 0xFFFFu;
        if (aligned && row0 + SS_WTILE <= n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int4 x = ld_stream(reinterpret_cast<const int4 *>(val + row0 + j * 128 + lane * 4));
                v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
            }
        } else {
            okmask = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t r = row0 + j * 128 + lane * 4 + k;
                    const bool ok = r < n;
                    okmask |= ok ? 1u << (4 * j + k) : 0u;
                    v[4 * j + k] = ok ? ld_stream(val + r) : 0;
                }
        }
        // survivors of the whole tile are ranked with ONE warp scan: the four 128-row groups'
        // counts (<= 128 each) ride in the four bytes of one word
        uint32_t need[4], packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t nd = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // bucket of d = v - lo, saturated at kSsBits: everything at or past the last
                // bound (and, by wrap-around, everything below the first) reads the zero word
                const uint32_t tb = min(((uint32_t)v[4 * j + k] - ulo) >> bsh, kSsBits);
                const uint32_t w = bits[tb >> 5];
                nd |= (__funnelshift_r(w, 0u, tb) & 1u) << k;             // shift count taken mod 32
            }
            nd &= okmask >> (4 * j);
            need[j] = nd;
            packed |= (uint32_t)__popc(nd) << (8 * j);
        }
        const uint32_t incl = warp_incl_scan(packed, lane);
        const uint32_t totp = __shfl_sync(kFull, incl, 31);
        if (totp == 0) continue;
        const uint32_t tot = __dp4a(totp, 0x01010101u, 0u);        // sum of the four group totals, <= 512
        if (wcount + tot > (uint32_t)SS_WORK) {                    // no room for this tile: resolve first
            __syncwarp();
            nhits = ss_resolve<PAIRS>(L, S.pack, plan.pair_depth, work, wcount, my_cnt, my_hits, nhits, lane, lt);
            wcount = 0;
            __syncwarp();
        }
        const uint32_t excl = incl - packed;
        uint32_t gbase = wcount;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t slot = gbase + ((excl >> (8 * j)) & 0xFFu);
            const uint32_t rel0 = t * SS_WTILE + j * 128 + lane * 4;     // row within the chunk
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (need[j] & (1u << k)) work[slot++] = make_uint2((uint32_t)v[4 * j + k], rel0 + k);
            gbase += (totp >> (8 * j)) & 0xFFu;
        }
        wcount += tot;
    }
    if (wcount) {
        __syncwarp();
        nhits = ss_resolve<PAIRS>(L, S.pack, plan.pair_depth, work, wcount, my_cnt, my_hits, nhits, lane, lt);
    }
    if (lane == 0) chunk_hits[chunk] = nhits;
    __syncwarp();
    // inclusive prefix over the interval counters (ids 0 .. m), in place
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 <= plan.m; i0 += kWarp) {
        const uint32_t i = i0 + lane;
        const uint32_t x = i <= plan.m ? my_cnt[i] : 0u;
        const uint32_t incl = warp_incl_scan(x, lane) + carry;
        if (i <= plan.m) my_cnt[i] = incl;
        carry = __shfl_sync(kFull, incl, 31);
    }
    __syncwarp();
    // query q covers intervals first[q] .. last[q] (ids; empty when first > last)
    for (uint32_t q = lane; q < plan.q_count; q += kWarp) {
        const uint32_t a = plan.q_first[q], b = plan.q_last[q];
        counts[(size_t)q * num_chunks + chunk] = a <= b ? my_cnt[b] - my_cnt[a - 1] : 0u;
    }
}

// One CTA per query: exclusive scan of its counts row in place, total to totals[q].
__global__ void __launch_bounds__(1024)
ss_offsets_kernel(uint32_t *__restrict__ counts, uint32_t num_chunks, int64_t *__restrict__ totals) {
    __shared__ uint32_t s_warp[32];
    uint32_t *row = counts + (size_t)blockIdx.x * num_chunks;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;                                    // identical in every thread
    for (uint32_t base = 0; base < num_chunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < num_chunks ? row[i] : 0u;
        const uint32_t incl = warp_incl_scan(x, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) s_warp[lane] = warp_incl_scan(s_warp[lane], lane);   // inclusive over warps
        __syncthreads();
        const uint32_t wexcl = warp ? s_warp[warp - 1] : 0u;
        if (i < num_chunks) row[i] = carry + wexcl + incl - x;
        carry += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = (int64_t)carry;
}

// Emit.  A warp walks its row range 32 rows at a time and turns every covered row into
// (query, row) pairs, row-major, in a small per-warp queue.  Whenever 32 pairs are parked
// they are written out in one go: lanes holding the same query find each other with one
// ballot per query-id bit, rank themselves in lane (= row) order and append behind that
// query's running offset.  Cost per pair is a handful of instructions whatever the number
// of distinct queries, where the previous version spent one reduce / ballot / update round
// per (row group, query).
__device__ __forceinline__ void ss_drain(uint32_t *queue, uint32_t head, uint32_t avail, uint32_t lane,
                                         uint32_t *run, uint32_t row_begin,
                                         int32_t *const *__restrict__ outs, int64_t capacity) {
    const bool live = lane < avail;
    uint32_t x = live ? queue[(head + lane) & (SS_QUEUE - 1)] : 0u;
    const uint32_t q = x >> 24;
    // (r01s: one MATCH.ANY instead of the eight ballots cut the kernel's instructions by a fifth
    // and made it slower, 110 -> 137 us: the instruction is that expensive on this part)
    const uint32_t peers = warp_match8(__ballot_sync(kFull, live), q);
    uint32_t old = 0;
    if (live) old = run[q];
    __syncwarp();
    if (live) {
        const uint32_t r = __popc(peers & ((1u << lane) - 1u));
        if (r == 0) run[q] = old + __popc(peers);
        const int64_t idx = (int64_t)old + r;
        if (idx < capacity) outs[q][idx] = (int32_t)(row_begin + (x & 0xFFFFFFu));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(SS_THREADS)
ss_emit_kernel(const uint32_t *__restrict__ hitlist, const uint32_t *__restrict__ chunk_hits,
               SharedScanPlan plan, uint32_t chunk_rows, uint32_t num_chunks,
               const uint32_t *__restrict__ offsets /* [q][num_chunks] */,
               int32_t *const *__restrict__ outs, int64_t capacity, uint32_t base_pos) {
    __shared__ uint16_t s_off[SS_BMAX + 2];
    __shared__ uint32_t s_run[SS_WARPS][SS_QMAX];
    __shared__ uint32_t s_queue[SS_WARPS][SS_QUEUE];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < plan.m + 2; i += SS_THREADS) s_off[i] = plan.cov_off[i];
    __syncthreads();
    const uint32_t chunk = blockIdx.x * SS_WARPS + warp;
    if (chunk >= num_chunks) return;
    const uint32_t nh = chunk_hits[chunk];
    if (nh == 0) return;
    uint32_t *run = s_run[warp];
    uint32_t *queue = s_queue[warp];
    for (uint32_t q = lane; q < plan.q_count; q += kWarp)
        run[q] = offsets[(size_t)q * num_chunks + chunk];
    __syncwarp();
    const uint32_t row_begin = chunk * chunk_rows + base_pos;      // only ever added to emitted rows
    const uint32_t *__restrict__ list = hitlist + (size_t)chunk * chunk_rows;
    const uint8_t *__restrict__ cov_q = plan.cov_q;
    uint32_t head = 0, avail = 0;                          // queue state, uniform across the warp
    uint32_t nxt = lane < nh ? list[lane] : 0u;
    for (uint32_t base = 0; base < nh; base += kWarp) {
        const uint32_t x = nxt;
        const bool live = base + lane < nh;
        if (base + kWarp + lane < nh) nxt = list[base + kWarp + lane];
        const uint32_t id = x >> 23, rel = x & 0x7FFFFFu;
        const uint32_t b = live ? s_off[id] : 0u, e = live ? s_off[id + 1] : 0u;
        const uint32_t ncov = e - b;                       // >= 1 for every listed row
        // Disjoint queries -- every listed row is wanted by exactly one -- need no queue: lanes
        // with the same query find each other (one ballot per query-id bit), rank themselves in
        // lane (= row) order and append behind that query's running offset.
        if (__all_sync(kFull, !live || ncov == 1u)) {
            if (avail) {                                   // pairs parked earlier come first
                ss_drain(queue, head, avail, lane, run, row_begin, outs, capacity);
                head += avail;
                avail = 0;
            }
            const uint32_t q = live ? (uint32_t)cov_q[b] : 0u;
            const uint32_t peers = warp_match8(__ballot_sync(kFull, live), q);
            uint32_t old = 0;
            if (live) old = run[q];
            __syncwarp();
            if (live) {
                const uint32_t r = __popc(peers & ((1u << lane) - 1u));
                if (r == 0) run[q] = old + __popc(peers);
                const int64_t idx = (int64_t)old + r;
                if (idx < capacity) outs[q][idx] = (int32_t)(row_begin + rel);
            }
            __syncwarp();
            continue;
        }
        const uint32_t incl = warp_incl_scan(ncov, lane);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        if (avail + total <= (uint32_t)SS_QUEUE) {
            uint32_t slot = head + avail + incl - ncov;
            for (uint32_t c = b; c < e; ++c, ++slot)
                queue[slot & (SS_QUEUE - 1)] = ((uint32_t)cov_q[c] << 24) | rel;
            avail += total;
            __syncwarp();
            while (avail >= kWarp) {
                ss_drain(queue, head, kWarp, lane, run, row_begin, outs, capacity);
                head += kWarp;
                avail -= kWarp;
            }
        } else {
            // heavily overlapping queries: more pairs than the queue holds.  Flush what is
            // parked, then serve these 32 rows query by query, smallest pending id first.
            if (avail) ss_drain(queue, head, avail, lane, run, row_begin, outs, capacity);
            head += avail;
            avail = 0;
            uint32_t c = b;
            while (true) {
                const uint32_t q = c < e ? (uint32_t)cov_q[c] : 0xFFFFFFFFu;
                const uint32_t qmin = __reduce_min_sync(kFull, q);
                if (qmin == 0xFFFFFFFFu) break;
                const bool mine = q == qmin;
                const uint32_t peers = __ballot_sync(kFull, mine);
                const uint32_t old = run[qmin];
                __syncwarp();
                if (lane == 0) run[qmin] = old + __popc(peers);
                __syncwarp();
                if (mine) {
                    const int64_t idx = (int64_t)old + __popc(peers & ((1u << lane) - 1u));
                    if (idx < capacity) outs[qmin][idx] = (int32_t)(row_begin + rel);
                    ++c;
                }
            }
        }
    }
    if (avail) ss_drain(queue, head, avail, lane, run, row_begin, outs, capacity);
}

// Emit over pair lists (SharedScanPlan::pair_depth > 0), sorted tile by tile (r01zg).  Every
// entry names its query, so no interval lookup and no pair queue.  r01zf tried the plain form
// first -- rank 32 entries with the ballots, store each at its query's cursor -- and ncu put a
// third of the stall samples on that store: 32 lanes, ~28 different queries, one LSU pass per
// touched sector, ten million of them, with every shared-memory access of the SM queueing
// behind (90-97 us).  Here a warp first sorts 1024 entries of its chunk's list by query inside
// shared memory (count, prefix over the queries, stable ranked scatter: the ballots rank, the
// scattered writes hit shared-memory banks instead of L2 sectors) and then writes the tile out
// in sorted order: 32 consecutive lanes cover three or four queries' runs, i.e. a handful of
// sectors per store instead of twenty-eight (72 us; the queue emit over interval lists: 105).
constexpr int SO_TILE = 1024;
constexpr int SO_QPAD = ((SS_QMAX + 31) / 32) * 32;          // 160

__global__ void __launch_bounds__(SS_THREADS)
ss_emit_sorted_kernel(const uint32_t *__restrict__ hitlist, const uint32_t *__restrict__ chunk_hits,
                      uint32_t q_count, uint32_t chunk_rows, uint32_t list_rows, uint32_t num_chunks,
                      const uint32_t *__restrict__ offsets /* [q][num_chunks] */,
                      int32_t *const *__restrict__ outs, int64_t capacity, uint32_t base_pos) {
    __shared__ uint32_t s_tile[SS_WARPS][SO_TILE];
    __shared__ uint32_t s_run[SS_WARPS][SO_QPAD];            // next free slot of every query's list
    __shared__ uint16_t s_toff[SS_WARPS][SO_QPAD];           // first tile slot of every query (< 1024)
    __shared__ uint32_t s_cur[SS_WARPS][SO_QPAD];            // counts, then scatter cursors
    __shared__ int32_t *s_out[SS_QMAX];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t q = threadIdx.x; q < q_count; q += SS_THREADS) s_out[q] = outs[q];
    __syncthreads();
    const uint32_t chunk = blockIdx.x * SS_WARPS + warp;
    if (chunk >= num_chunks) return;
    const uint32_t nh = chunk_hits[chunk];
    if (nh == 0) return;
    uint32_t *tile = s_tile[warp], *run = s_run[warp], *cur = s_cur[warp];
    uint16_t *toff = s_toff[warp];
    for (uint32_t q = lane; q < SO_QPAD; q += kWarp)
        run[q] = q < q_count ? offsets[(size_t)q * num_chunks + chunk] : 0u;
    const uint32_t row_begin = chunk * chunk_rows + base_pos;      // only ever added to emitted rows
    const uint32_t *__restrict__ list = hitlist + (size_t)chunk * list_rows;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t q_pad = ((q_count + kWarp - 1) / kWarp) * kWarp;
    for (uint32_t t0 = 0; t0 < nh; t0 += SO_TILE) {
        const uint32_t tn = min((uint32_t)SO_TILE, nh - t0);
        for (uint32_t q = lane; q < q_pad; q += kWarp) cur[q] = 0;
        __syncwarp();
        // (a) entries per query
        for (uint32_t i = lane; i < tn; i += kWarp) atomicAdd(&cur[list[t0 + i] >> 23], 1u);
        __syncwarp();
        // (b) exclusive prefix over the queries -> first tile slot; the counts become cursors
        uint32_t carry = 0;
        for (uint32_t q0 = 0; q0 < q_pad; q0 += kWarp) {
            const uint32_t c = cur[q0 + lane];
            const uint32_t incl = warp_incl_scan(c, lane) + carry;
            toff[q0 + lane] = (uint16_t)(incl - c);
            cur[q0 + lane] = incl - c;
            carry = __shfl_sync(kFull, incl, 31);
        }
        __syncwarp();
        // (c) stable scatter into the tile: batches in list order, lanes of one query ranked in
        //     lane order behind that query's cursor
        uint32_t nxt = lane < tn ? list[t0 + lane] : 0u;
        for (uint32_t i0 = 0; i0 < tn; i0 += kWarp) {
            const uint32_t x = nxt;
            const bool live = i0 + lane < tn;
            if (i0 + kWarp + lane < tn) nxt = list[t0 + i0 + kWarp + lane];
            const uint32_t q = live ? x >> 23 : 0u;
            const uint32_t peers = warp_match8(__ballot_sync(kFull, live), q);
            uint32_t old = 0;
            if (live) old = cur[q];
            __syncwarp();
            if (live) {
                const uint32_t r = __popc(peers & lt);
                if (r == 0) cur[q] = old + __popc(peers);
                tile[old + r] = x;
            }
            __syncwarp();
        }
        // (d) write the tile out in sorted order
        for (uint32_t sl = lane; sl < tn; sl += kWarp) {
            const uint32_t x = tile[sl], q = x >> 23;
            const int64_t o = (int64_t)run[q] + (sl - toff[q]);
            if (o < capacity) s_out[q][o] = (int32_t)(row_begin + (x & 0x7FFFFFu));
        }
        __syncwarp();
        // (e) advance the lists: cursor - first slot = entries of this tile
        for (uint32_t q = lane; q < q_pad; q += kWarp) run[q] += cur[q] - toff[q];
        __syncwarp();
    }
}

// r01v-y, tried and dropped: a lane-independent emit (every lane owns a contiguous slice of the
// chunk's hit list, counts per query into its own cell of a [query][lane] matrix in shared
// memory, prefixes over the lanes, second walk writes at base[q] + prefix++).  It needs 35 M warp
// instructions against the 59 M of the ballot form above and was slower in every variant tried
// (4-byte loads 158 us, 16-byte loads 140 us, fire-and-forget shared atomics + prefetch 146 us,
// against 110 us): the per-lane walk of the list is a chain of dependent L2 / DRAM loads at
// 25 % occupancy (ncu r01x: issue active 26 %, 6.6 warps per issue waiting on global loads),
// where the ballot form reads the list coalesced, one line per 32 hits.

// ---- launch -------------------------------------------------------------------------------------
SharedScanGeom shared_scan_geom(uint32_t n, int sm_count) {
    SharedScanGeom g{};
    const uint32_t wtiles = (n + SS_WTILE - 1) / SS_WTILE;
    const uint32_t max_chunks = (uint32_t)sm_count * 4u * SS_WARPS;
    uint32_t tiles_per_chunk = (wtiles + max_chunks - 1) / max_chunks;
    if (tiles_per_chunk == 0) tiles_per_chunk = 1;
    g.chunk_rows = tiles_per_chunk * SS_WTILE;
    g.num_chunks = (wtiles + tiles_per_chunk - 1) / tiles_per_chunk;
    g.grid = (g.num_chunks + SS_WARPS - 1) / SS_WARPS;
    return g;
}

int launch_shared_classify(const int32_t *val, uint32_t n, const SharedScanPlan &plan,
                           const SharedScanGeom &g, uint32_t *hitlist, uint32_t *chunk_hits,
                           uint32_t *counts, int64_t *totals, cudaStream_t s) {
    // function attributes are per device: one flag per device the engine has launched on
    static bool attr_set_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool &attr_set = attr_set_dev[dev & 63];
    if (!attr_set) {
        cudaFuncSetAttribute(ss_classify_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SsShared));
        cudaFuncSetAttribute(ss_classify_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SsShared));
        attr_set = true;
    }
    if (plan.pair_depth)
        ss_classify_kernel<true><<<g.grid, SS_THREADS, sizeof(SsShared), s>>>(val, n, plan, g.chunk_rows, g.num_chunks,
                                                                              hitlist, chunk_hits, counts);
    else
        ss_classify_kernel<false><<<g.grid, SS_THREADS, sizeof(SsShared), s>>>(val, n, plan, g.chunk_rows, g.num_chunks,
                                                                               hitlist, chunk_hits, counts);
    ss_offsets_kernel<<<plan.q_count, 1024, 0, s>>>(counts, g.num_chunks, totals);
    return 2;
}

int launch_shared_emit(const uint32_t *hitlist, const uint32_t *chunk_hits,
                       const SharedScanPlan &plan, const SharedScanGeom &g, const uint32_t *offsets,
                       int32_t *const *outs, int64_t capacity, uint32_t base_pos, cudaStream_t s) {
    if (plan.pair_depth)
        ss_emit_sorted_kernel<<<g.grid, SS_THREADS, 0, s>>>(hitlist, chunk_hits, plan.q_count, g.chunk_rows,
                                                            g.chunk_rows * plan.pair_depth, g.num_chunks, offsets,
                                                            outs, capacity, base_pos);
    else
        ss_emit_kernel<<<g.grid, SS_THREADS, 0, s>>>(hitlist, chunk_hits, plan, g.chunk_rows,
                                                     g.num_chunks, offsets, outs, capacity, base_pos);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_shared_scan() {
    { auto *fp = &ss_classify_kernel<true>; preload_one(reinterpret_cast<const void *>(fp)); }
    { auto *fp = &ss_classify_kernel<false>; preload_one(reinterpret_cast<const void *>(fp)); }
    preload_one(reinterpret_cast<const void *>(&ss_emit_kernel));
    preload_one(reinterpret_cast<const void *>(&ss_emit_sorted_kernel));
    preload_one(reinterpret_cast<const void *>(&ss_offsets_kernel));
}

}  // namespace adb
