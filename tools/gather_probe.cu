// gather_probe.cu -- which load flavour avoids L2 over-fetch for sparse gathers? (dev tool)
// out[i] = col[pos[i]] with pos ascending at density d.  Prints time per variant; run under
// ncu --metrics dram__bytes_read.sum to see the traffic.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

template <int V> __device__ __forceinline__ int ld(const int* p);
template <> __device__ __forceinline__ int ld<0>(const int* p) { return __ldg(p); }
template <> __device__ __forceinline__ int ld<1>(const int* p) { int r; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<2>(const int* p) { int r; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<3>(const int* p) { int r; asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<4>(const int* p) { int r; asm volatile("ld.global.lu.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<5>(const int* p) { int r; asm volatile("ld.global.nc.L1::evict_last.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<6>(const int* p) { int r; asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }

template <> __device__ __forceinline__ int ld<7>(const int* p) { int r; asm volatile("ld.global.L2::64B.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<8>(const int* p) { int r; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<9>(const int* p) {
    int r; unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r; }
template <> __device__ __forceinline__ int ld<10>(const int* p) { int r; asm volatile("ld.global.nc.L1::no_allocate.L2::256B.s32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
template <> __device__ __forceinline__ int ld<11>(const int* p) {     // 1-byte load of the wanted word's low byte
    unsigned r; asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p)); return (int)r; }

template <int V>
__global__ void __launch_bounds__(256) gather(const int* __restrict__ col, const int* __restrict__ pos, long n, int* __restrict__ out) {
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = ld<V>(col + pos[i]);
}
__global__ void fill(int* col, long n) { long s=(long)gridDim.x*blockDim.x; for (long i=(long)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=s) col[i]=(int)(i*2654435761u); }
__global__ void mkpos(int* pos, long h, long step) { long s=(long)gridDim.x*blockDim.x; for (long i=(long)blockIdx.x*blockDim.x+threadIdx.x;i<h;i+=s) pos[i]=(int)(i*step + ((i*2654435761u)>>8)%step); }

template <int V> void run(const char* name, const int* col, const int* pos, long h, int* out) {
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<V><<<148*8,256>>>(col,pos,h,out);
    cudaEventRecord(a);
    for (int r=0;r<5;++r) gather<V><<<148*8,256>>>(col,pos,h,out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b);
    printf("%-28s %8.1f us  (%.1f Ggather/s)\n", name, ms/5*1e3, h/(ms/5*1e-3)/1e9);
}
int main(int argc, char** argv) {
    long n = 500000000, step = argc>1?atol(argv[1]):100; long h = n/step;
    if (argc>2) { size_t g=atoi(argv[2]); cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,g); size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity); printf("set L2 fetch granularity %zu -> %s, now %zu\n", g, cudaGetErrorString(e), got);}
    else { size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", got);}
    int *col,*pos,*out; cudaMalloc(&col,n*4); cudaMalloc(&pos,h*4); cudaMalloc(&out,h*4);
    fill<<<148*8,256>>>(col,n); mkpos<<<148*8,256>>>(pos,h,step); cudaDeviceSynchronize();
    printf("n=%ld hits=%ld (1 per %ld rows)\n", n,h,step);
    run<0>("__ldg (ld.global.nc)",col,pos,h,out);
    run<1>("nc.L1::no_allocate",col,pos,h,out);
    run<2>("ld.global.cg",col,pos,h,out);
    run<3>("ld.global.cs",col,pos,h,out);
    run<4>("ld.global.lu",col,pos,h,out);
    run<5>("nc.L1::evict_last",col,pos,h,out);
    run<6>("ld.volatile",col,pos,h,out);
    run<7>("ld.global.L2::64B",col,pos,h,out);
    run<8>("nc.no_allocate.L2::64B",col,pos,h,out);
    run<9>("nc.no_allocate.L2 evict_first hint",col,pos,h,out);
    run<10>("nc.no_allocate.L2::256B",col,pos,h,out);
    run<11>("nc.no_allocate.u8",col,pos,h,out);
    return 0;
}
