// random_read_probe.cu -- how many random 16-byte reads per second does this part serve, as a
// function of the table size (L2-resident ... DRAM) and of the miss granularity?  (dev tool)
//
// The hash-join probe (hash_join.cu: hj_probe_kernel) is one random 16-byte slot read per probe
// row into a ~2 GB table and runs at 31 G reads/s whatever the miss size;
// this probe isolates that access pattern: out[i] = table[hash(i) & mask].x, 100 M reads,
// tables of 32 MB ... 4 GB, plain loads vs ld.global.nc.L2::64B, and a "windowed" order in
// which consecutive warps stay inside one 32 MB slice of the table (what a probe side
// partitioned by the top hash bits would see).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/random_read_probe tools/random_read_probe.cu
//   tools/build/random_read_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint4 ld_plain(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint4 ld_64(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// window_slots = 0: every read anywhere in the table; otherwise read i stays inside the
// window_slots-slot slice number (i / reads_per_window)
template <bool L64>
__global__ void __launch_bounds__(256)
probe(const uint4 *__restrict__ table, uint32_t slot_mask, uint32_t n, uint32_t window_slots,
      uint32_t reads_per_window, uint32_t *__restrict__ out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t s = mix(i) & slot_mask;
        if (window_slots) s = ((i / reads_per_window) * window_slots + (s & (window_slots - 1))) & slot_mask;
        const uint4 v = L64 ? ld_64(table + s) : ld_plain(table + s);
        out[i] = v.x + v.w;
    }
}

__global__ void fill(uint4 *t, size_t slots) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += stride)
        t[i] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}

template <bool L64>
static void run(const char *name, const uint4 *table, uint32_t slot_mask, uint32_t n, uint32_t window_slots,
                uint32_t *out) {
    const uint32_t windows = window_slots ? (slot_mask + 1) / window_slots : 1;
    const uint32_t rpw = window_slots ? (n + windows - 1) / windows : n;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    probe<L64><<<148 * 8, 256>>>(table, slot_mask, n, window_slots, rpw, out);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) probe<L64><<<148 * 8, 256>>>(table, slot_mask, n, window_slots, rpw, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("  %-34s %8.3f ms  %6.1f G reads/s\n", name, ms / 3, n / (ms / 3 * 1e-3) / 1e9);
}

int main() {
    const uint32_t n = 100000000;
    uint32_t *out;
    cudaMalloc(&out, (size_t)n * 4);
    for (int lg = 21; lg <= 28; ++lg) {                       // 2^21 .. 2^28 slots of 16 bytes: 32 MB .. 4 GB
        const size_t slots = (size_t)1 << lg;
        uint4 *table;
        if (cudaMalloc(&table, slots * 16) != cudaSuccess) break;
        fill<<<148 * 8, 256>>>(table, slots);
        cudaDeviceSynchronize();
        printf("table %6zu MB, %u random 16-byte reads\n", slots * 16 >> 20, n);
        run<false>("plain load (128-byte misses)", table, (uint32_t)slots - 1, n, 0, out);
        run<true>("ld.global.nc.L2::64B", table, (uint32_t)slots - 1, n, 0, out);
        if (lg > 21) run<true>("L2::64B, 32 MB windows in order", table, (uint32_t)slots - 1, n, 1u << 21, out);
        cudaFree(table);
    }
    cudaFree(out);
    return 0;
}
