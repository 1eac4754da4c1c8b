// Experiment (not product code): streaming copy with 128-bit vs 256-bit global accesses on sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint4 ld16(const void *p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ void st16(void *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
struct u8 { unsigned v[8]; };
__device__ __forceinline__ u8 ld32(const void *p, int hint) {
    u8 r;
    if (hint)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st32(void *p, const u8 &r) {
    asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
}
template <int U>
__global__ void __launch_bounds__(128) copy128(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
    size_t i = ((size_t)blockIdx.x * U) * 128 + threadIdx.x;
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 128 < n) r[u] = ld16(src + i + u * 128);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 128 < n) st16(dst + i + u * 128, r[u]);
}
template <int U, int HINT>
__global__ void __launch_bounds__(128) copy256(const u8 *__restrict__ src, u8 *__restrict__ dst, size_t n) {
    size_t i = ((size_t)blockIdx.x * U) * 128 + threadIdx.x;
    u8 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 128 < n) r[u] = ld32(src + i + u * 128, HINT);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * 128 < n) st32(dst + i + u * 128, r[u]);
}
extern "C" int run_copy(int mode, const void *src, void *dst, size_t bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) { size_t n = bytes / 16; copy128<4><<<(unsigned)((n + 511) / 512), 128, 0, st>>>((const uint4 *)src, (uint4 *)dst, n); }
    if (mode == 1) { size_t n = bytes / 16; copy128<8><<<(unsigned)((n + 1023) / 1024), 128, 0, st>>>((const uint4 *)src, (uint4 *)dst, n); }
    if (mode == 2) { size_t n = bytes / 32; copy256<2, 0><<<(unsigned)((n + 255) / 256), 128, 0, st>>>((const u8 *)src, (u8 *)dst, n); }
    if (mode == 3) { size_t n = bytes / 32; copy256<4, 0><<<(unsigned)((n + 511) / 512), 128, 0, st>>>((const u8 *)src, (u8 *)dst, n); }
    if (mode == 4) { size_t n = bytes / 32; copy256<4, 1><<<(unsigned)((n + 511) / 512), 128, 0, st>>>((const u8 *)src, (u8 *)dst, n); }
    return (int)cudaGetLastError();
}
