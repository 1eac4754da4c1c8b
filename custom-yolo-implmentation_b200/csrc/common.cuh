// Shared device/host helpers for the box-path kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include "../../include/yolo_boxpath.h"

namespace yb {

// ---- host-side error plumbing (cabi.cu) -------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void note_launch();

#define YB_CUDA(call)                                            \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return ::yb::cuda_fail(e__, #call); \
    } while (0)

// after every kernel launch: count it (yb_launch_count) and surface a launch error
#define YB_LAUNCH_CHECK()                  \
    do {                                   \
        ::yb::note_launch();               \
        YB_CUDA(cudaGetLastError());       \
    } while (0)
#define YB_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            ::yb::set_error(__VA_ARGS__); \
            return YB_ERR_ARG;           \
        }                                \
    } while (0)

// An NVTX range around each entry point of the ABI (the reference wraps nothing: SURVEY.md §5 asks for the ranges a
// timeline needs to attribute the launches to the call that issued them).  Header-only NVTX3: a no-op costing one
// indirect call unless a tool (nsys, ncu --nvtx) is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define YB_NVTX(name) ::yb::NvtxRange nvtx_range__(name)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

constexpr int kRegMax = 16;        // bins per box side; the only value the kernels are built for
constexpr float kEpsIou = 1e-6f;   // src/model/losses.py:40
constexpr float kEpsLog = 1e-12f;  // src/model/losses.py:53-54

// ---- 128-bit streaming access --------------------------------------------------------------
// Head outputs and gradients are touched exactly once per step: loads bypass L1 allocation and
// stores use the streaming (evict-first) policy so the small reused tables (GT, match table,
// anchors) stay resident in L1/L2.
__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
    uint4 r;
    // not volatile: a pure read of data no kernel writes, so the compiler may hoist / batch it
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
        : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream16(void *p, const uint4 &v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

template <typename T>
struct ElemsPer16;
template <>
struct ElemsPer16<float> {
    static constexpr int value = 4;
};
template <>
struct ElemsPer16<__nv_bfloat16> {
    static constexpr int value = 8;
};

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // round-to-nearest-even, as Tensor.to(bfloat16)
    return *reinterpret_cast<uint32_t *>(&v);
}

// A "group" is the VW consecutive anchors one thread owns: VW = 16 B worth of elements on the
// vector path, 1 on the scalar fall-back (row pitch or base pointer not 16-byte aligned).
template <typename T, int VW>
struct Group;

template <>
struct Group<float, 4> {
    uint4 raw;
    __device__ __forceinline__ void load(const float *p) { raw = ldg_stream16(p); }
    __device__ __forceinline__ float get(int v) const {
        return __uint_as_float(v == 0 ? raw.x : v == 1 ? raw.y : v == 2 ? raw.z : raw.w);
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[4]) {
        stg_stream16(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                                   __float_as_uint(f[3])));
    }
    static __device__ __forceinline__ void store_zero(float *p) { stg_stream16(p, make_uint4(0, 0, 0, 0)); }
};

template <>
struct Group<__nv_bfloat16, 8> {
    uint4 raw;
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { raw = ldg_stream16(p); }
    __device__ __forceinline__ float get(int v) const {
        uint32_t w = (v >> 1) == 0 ? raw.x : (v >> 1) == 1 ? raw.y : (v >> 1) == 2 ? raw.z : raw.w;
        return (v & 1) ? bf16_hi(w) : bf16_lo(w);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[8]) {
        stg_stream16(p, make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                   pack_bf16x2(f[6], f[7])));
    }
    static __device__ __forceinline__ void store_zero(__nv_bfloat16 *p) { stg_stream16(p, make_uint4(0, 0, 0, 0)); }
};

template <>
struct Group<float, 1> {
    float raw;
    __device__ __forceinline__ void load(const float *p) { raw = __ldg(p); }
    __device__ __forceinline__ float get(int) const { return raw; }
    static __device__ __forceinline__ void store(float *p, const float (&f)[1]) { *p = f[0]; }
    static __device__ __forceinline__ void store_zero(float *p) { *p = 0.f; }
};

template <>
struct Group<__nv_bfloat16, 1> {
    __nv_bfloat16 raw;
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { raw = *p; }
    __device__ __forceinline__ float get(int) const { return __bfloat162float(raw); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[1]) { *p = __float2bfloat16_rn(f[0]); }
    static __device__ __forceinline__ void store_zero(__nv_bfloat16 *p) { *p = __float2bfloat16_rn(0.f); }
};

template <typename T>
__device__ __forceinline__ float load_as_float(const T *p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void store_from_float(T *p, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// ---- DFL softmax-expectation of one box side ------------------------------------------------
// softmax over the 16 bins followed by sum(p * [0..15]) (src/model/losses.py:157-159), evaluated as
// (sum_j j e_j) / (sum_j e_j) with e_j = exp(x_j - max): one division per side instead of sixteen.
// ex2.approx keeps e_j within ~2^-21 relative of exp(); the result is within ~1e-6 relative of the
// reference's, far inside the 1e-5 budget (the matched-anchor decision is re-checked against the
// oracle with its runner-up margin in tests/test_gpu_loss.py).
__device__ __forceinline__ float dfl_expectation16(const float (&x)[16], float (&prob)[16], bool want_prob = false) {
    float m = x[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) m = fmaxf(m, x[j]);
    float e[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = __expf(x[j] - m);
    float s8[8], s4[4], w8[8], w4[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s8[j] = e[j] + e[j + 8];
        w8[j] = fmaf((float)(j + 8), e[j + 8], (float)j * e[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s4[j] = s8[j] + s8[j + 4];
        w4[j] = w8[j] + w8[j + 4];
    }
    const float sum = (s4[0] + s4[2]) + (s4[1] + s4[3]);
    const float wsum = (w4[0] + w4[2]) + (w4[1] + w4[3]);
    if (want_prob) {
        const float inv = __fdividef(1.f, sum);
#pragma unroll
        for (int j = 0; j < 16; ++j) prob[j] = e[j] * inv;
    }
    return __fdiv_rn(wsum, sum);
}

// Pixel-space box of one anchor from its four expected distances (src/model/losses.py:178-186).
struct PredBox {
    float x1, y1, x2, y2, cx, cy, w, h;
};
__device__ __forceinline__ PredBox decode_box(float ax, float ay, float s, float dl, float dt, float dr, float db) {
    PredBox b;
    b.x1 = __fmul_rn(__fsub_rn(ax, dl), s);
    b.y1 = __fmul_rn(__fsub_rn(ay, dt), s);
    b.x2 = __fmul_rn(__fadd_rn(ax, dr), s);
    b.y2 = __fmul_rn(__fadd_rn(ay, db), s);
    b.w = __fsub_rn(b.x2, b.x1);
    b.h = __fsub_rn(b.y2, b.y1);
    b.cx = __fmul_rn(__fadd_rn(b.x1, b.x2), 0.5f);
    b.cy = __fmul_rn(__fadd_rn(b.y1, b.y2), 0.5f);
    return b;
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): two IEEE round-to-nearest operations per
// instruction, lane-wise identical to the scalar __fmul_rn / __fmaf_rn / __fadd_rn ------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ float fast_ex2(float x) {         // MUFU.EX2, 2 ulp
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float fast_rcp(float x) {        // MUFU.RCP, 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The same expectation from two 8-bin halves (online-softmax merge): lets a kernel keep only 8 rows
// of a side in registers at a time.  Half h holds bins 8h .. 8h+7.  exp(x - max) is evaluated in the
// base-2 domain, ex2(x * log2e - o) with o = fl(max * log2e): packed FFMA2 (one instruction per two
// bins) + MUFU.EX2.  The partial carries o itself, so the merge rescales by ex2(o_a - o) exactly as the
// terms were scaled and the rounding of o cancels.  Sums run on the packed pipe in the fixed order
// ((0+4)+(2+6)) + ((1+5)+(3+7)).
struct DflPartial {
    float o, s, w;      // o = fl(max * log2e), s = sum ex2(x*log2e - o), w = sum j ex2(x*log2e - o)
};
__device__ __forceinline__ DflPartial dfl_half8(const float (&x)[8], int bin0) {
    constexpr float kLog2e = 1.4426950408889634f;
    DflPartial r;
    const float m = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
    r.o = m * kLog2e;
    const f32x2 no2 = pack2(-r.o, -r.o), l2 = pack2(kLog2e, kLog2e);
    f32x2 e2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float lo, hi;
        unpack2(fma2(pack2(x[2 * k], x[2 * k + 1]), l2, no2), lo, hi);
        e2[k] = pack2(fast_ex2(lo), fast_ex2(hi));
    }
    float lo, hi;
    unpack2(add2(add2(e2[0], e2[2]), add2(e2[1], e2[3])), lo, hi);
    r.s = lo + hi;
    const float b = (float)bin0;
    f32x2 w2 = mul2(pack2(b + 6.f, b + 7.f), e2[3]);
    w2 = fma2(pack2(b + 4.f, b + 5.f), e2[2], w2);
    w2 = fma2(pack2(b + 2.f, b + 3.f), e2[1], w2);
    w2 = fma2(pack2(b, b + 1.f), e2[0], w2);
    unpack2(w2, lo, hi);
    r.w = lo + hi;
    return r;
}
__device__ __forceinline__ float dfl_merge(const DflPartial &a, const DflPartial &b) {
    const float o = fmaxf(a.o, b.o);
    const float fa = fast_ex2(a.o - o), fb = fast_ex2(b.o - o);
    return __fdiv_rn(fmaf(a.w, fa, b.w * fb), fmaf(a.s, fa, b.s * fb));
}

// log(1 + f) for two values f in [-0.293, 0] on the packed pipe: f + f^2 R(f), R a degree-5 minimax
// polynomial (max relative error 9.4e-8, fitted offline; the same coefficients as csrc/loss.cu's QFL path)
__device__ __forceinline__ f32x2 log1p_neg_small2(f32x2 f) {
    f32x2 r = pack2(0.3410167098045349f, 0.3410167098045349f);
    r = fma2(r, f, pack2(-0.08926734328269958f, -0.08926734328269958f));
    r = fma2(r, f, pack2(0.21280372142791748f, 0.21280372142791748f));
    r = fma2(r, f, pack2(-0.249073788523674f, -0.249073788523674f));
    r = fma2(r, f, pack2(0.33335742354393005f, 0.33335742354393005f));
    r = fma2(r, f, pack2(-0.49999991059303284f, -0.49999991059303284f));
    return fma2(mul2(f, f), r, f);
}

// ---- in-kernel dependencies between the CTAs of ONE launch --------------------------------------
// A launch may hold CTAs of several roles where a later role consumes what an earlier one produced (the matched-anchor
// terms need every tile of their image; the per-GT top-k needs the image's decoded boxes).  Producers count themselves
// off on a per-image counter when their global writes are done; a consumer polls that counter before it starts.
// Consumers always sit LATER in the grid's linear block order than everything they wait for, and the hardware hands out
// the blocks of a one-dimensional grid in that order, so whatever a resident consumer waits for is resident or finished:
// no deadlock.  That dispatch order is an observation, not a documented guarantee, hence the bounded poll: a consumer
// that waits longer than ~2 s raises *timeout_flag (reported through the entry point's out_loss and raised by the host
// side) and carries on, so a violated assumption shows up as an error, never as a hang.
#ifndef YB_DEP_MAX_SLEEP_NS            // longest sleep between two polls of a waiting consumer
#define YB_DEP_MAX_SLEEP_NS 1024
#endif
__device__ __forceinline__ void dep_signal(unsigned int *counter) {      // ONE thread, after a __syncthreads()
#ifndef YB_DEP_NOFENCE                                     // (measurement aid: what the fence costs)
    __threadfence();                                      // the CTA's writes (ordered before by the barrier) are visible first
#endif
    atomicAdd(counter, 1u);
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dep_wait(const unsigned int *counter, unsigned int target, unsigned int *timeout_flag,
                                         unsigned int max_sleep_ns = YB_DEP_MAX_SLEEP_NS) {
    if (ld_acquire_u32(counter) >= target) return;
    const long long t0 = clock64();
    unsigned int ns = 64;
    while (ld_acquire_u32(counter) < target) {
        __nanosleep(ns);
        ns = min(ns * 2u, max_sleep_ns);
        if (clock64() - t0 > (4ll << 30)) {               // ~2 s at 2 GHz
            atomicOr(timeout_flag, 1u);
            return;
        }
    }
#ifdef YB_DEP_STATS                                        // (measurement aid: how long consumers wait; timeout_flag[5], [6])
    atomicAdd(timeout_flag + 5, 1u);
    atomicAdd(timeout_flag + 6, (unsigned int)((clock64() - t0) >> 10));
#endif
}

// ---- programmatic dependent launch: the next kernel of the stream becomes resident while this one drains ----------
// A kernel launched through launch_pdl may start before its predecessor in the stream has finished; it must call
// pdl_wait() before it touches anything the predecessor wrote (or writes anything the predecessor reads).  The
// predecessor calls pdl_launch_dependents() once all of ITS blocks are allowed to be joined by the successor's -- here
// at its very start, so the successor's blocks fill the SMs its last wave leaves idle and wait there.
// A coherent (not .nc) scalar load that the compiler may still schedule freely: for data a PREDECESSOR kernel wrote, read
// by a kernel launched under programmatic dependent launch.  Its address must depend on something loaded after
// pdl_wait() -- that dependence, not a memory clobber, is what keeps it behind the wait.
__device__ __forceinline__ float ld_dependent_f32(const float *p) {
    float r;
    asm("ld.global.ca.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double warp_sum_d(double v) {      // butterfly: a fixed tree, every lane ends with the same sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace yb
