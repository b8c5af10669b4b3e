// TEST INFRASTRUCTURE — host-side SIMT simulator for the kernels in ds-gcn_b200/csrc.
//
// Compiled only with -DDSG_EMU by tests/emu/build_emu.py (g++), never by the product build
// and never loaded by the product package.  It lets the *same kernel sources* run on host
// memory so index logic, prologue/epilogue algebra and the C++ orchestration can be checked
// against the oracle in this GPU-less container.  One ucontext fiber per CUDA thread; the
// blocks of a launch run one after another (so a function-local `static` array is a faithful
// stand-in for __shared__); __syncthreads / warp shuffles are cooperative yields.
#pragma once
#include <ucontext.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

struct bf16 { uint16_t v; };
static inline bf16 __float2bfloat16(float f) {
    unsigned u = __float_as_uint(f);
    bf16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.v = 0x7fff; return r; }
    u += 0x7fffu + ((u >> 16) & 1u);   // round to nearest even
    r.v = (uint16_t)(u >> 16);
    return r;
}
static inline float __bfloat162float(bf16 b) { return __uint_as_float(((unsigned)b.v) << 16); }

typedef void* dsg_stream_t;

namespace emu {
struct Fiber {
    ucontext_t ctx;
    dim3 tid;
    bool done;
    char* stack;
};
struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber* cur = nullptr;
    dim3 bidx, bdim, gdim;
    int n_alive = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    // warp exchange
    std::vector<float> wbuf;       // [warp][32]
    std::vector<int> warrive;      // [warp]
    std::vector<unsigned> wgen;    // [warp]
    unsigned char* dyn_smem = nullptr;
    size_t dyn_cap = 0;
    std::function<void()> body;
    const char* error = nullptr;
};
State& st();
void yield();
void syncthreads();
float shfl(float v, int src_lane_fn_kind, int arg);   // kind 0: xor, 1: idx, 2: down
void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body);
}  // namespace emu

#define threadIdx (emu::st().cur->tid)
#define blockIdx (emu::st().bidx)
#define blockDim (emu::st().bdim)
#define gridDim (emu::st().gdim)
#define __syncthreads() emu::syncthreads()
#define DSG_DYN_SMEM(name) unsigned char* name = emu::st().dyn_smem
#define DSG_SHARED static
#define DSG_SET_SMEM(kernel, bytes) ((void)0)

static inline float __shfl_xor_sync(unsigned, float v, int m) { return emu::shfl(v, 0, m); }
static inline float __shfl_sync(unsigned, float v, int l) { return emu::shfl(v, 1, l); }
static inline float __shfl_down_sync(unsigned, float v, int d) { return emu::shfl(v, 2, d); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)__float_as_uint(emu::shfl(__uint_as_float((unsigned)v), 0, m)); }
static inline int __shfl_sync(unsigned, int v, int l) { return (int)__float_as_uint(emu::shfl(__uint_as_float((unsigned)v), 1, l)); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::shfl(0.f, 1, 0); }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }

template <class K, class... A>
static inline void dsg_launch(K kernel, dim3 grid, dim3 block, size_t smem, dsg_stream_t, A... args) {
    emu::launch(grid, block, smem, [=]() { kernel(args...); });
}
static inline const char* dsg_launch_error() {
    const char* e = emu::st().error;
    emu::st().error = nullptr;
    return e;
}
