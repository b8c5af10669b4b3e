// Device-code vocabulary shared by every kernel in this library.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a  (DSG_EMU undefined).
// Test build   : g++ -DDSG_EMU  -> the same kernel sources run on a host-side SIMT
//                simulator (tests/emu/emu_runtime.h: one fiber per CUDA thread, blocks run
//                one after another).  That build is TEST INFRASTRUCTURE for developing the
//                kernels' index logic without a GPU; the product package never loads it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "dsgcn_b200.h"

#ifdef DSG_EMU
#include "emu_runtime.h"
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define DSG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define DSG_SHARED __shared__
typedef __nv_bfloat16 bf16;
typedef cudaStream_t dsg_stream_t;
template <class K, class... A>
static inline void dsg_launch(K kernel, dim3 grid, dim3 block, size_t smem, dsg_stream_t st, A... args) {
    kernel<<<grid, block, smem, st>>>(args...);
}
#define DSG_SET_SMEM(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
static inline const char* dsg_launch_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
#endif

// SM count of the current device (grid sizing: one wave of resident CTAs); 148 on the B200, queried once so other Blackwell SKUs fill too
static inline int dsg_num_sms() {
#ifdef DSG_EMU
    return 148;
#else
    static int n = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
#endif
}

#define DSG_HD __host__ __device__ __forceinline__
#define DSG_D __device__ __forceinline__

// ---- scalar load/store with conversion ------------------------------------------------
template <class T> DSG_D float ldf(const T* p);
template <> DSG_D float ldf<float>(const float* p) { return *p; }
template <> DSG_D float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <class T> DSG_D void stf(T* p, float v);
template <> DSG_D void stf<float>(float* p, float v) { *p = v; }
template <> DSG_D void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// 4 consecutive elements (pointer must be 16B / 8B aligned)
template <class T> DSG_D float4 ldf4(const T* p);
template <> DSG_D float4 ldf4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> DSG_D float4 ldf4<bf16>(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float4 r;
    r.x = __uint_as_float(u.x << 16); r.y = __uint_as_float(u.x & 0xffff0000u);
    r.z = __uint_as_float(u.y << 16); r.w = __uint_as_float(u.y & 0xffff0000u);
    return r;
}
template <class T> DSG_D void stf4(T* p, float4 v);
template <> DSG_D void stf4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
DSG_D uint32_t dsg_pack_bf16x2(float lo, float hi) {
#ifndef DSG_EMU
    uint32_t r;                                                    // one packed conversion (round-to-nearest-even both halves)
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
#else
    bf16 a = __float2bfloat16(lo), b = __float2bfloat16(hi);
    return (uint32_t)(*reinterpret_cast<uint16_t*>(&a)) | ((uint32_t)(*reinterpret_cast<uint16_t*>(&b)) << 16);
#endif
}
template <> DSG_D void stf4<bf16>(bf16* p, float4 v) {
    uint2 u;
    u.x = dsg_pack_bf16x2(v.x, v.y);
    u.y = dsg_pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

DSG_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DSG_D float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- "activation source": a logical activation tensor defined on the fly --------------
//   value(r,c) = f( a1[c]*x1[r*ld1+c] + b1[c] + a2[c]*x2[r*ld2+c] + b2[c] ),  f = relu or identity
// (struct dsg_act_src in include/dsgcn_b200.h).  This one form covers BN-apply(+ReLU) of a raw conv
// output, "relu(bn(z) + residual)" and the BN-backward combination dy = ca*e + cb*y + cc.
typedef dsg_act_src ActSrc;

template <class T> DSG_D float act_value(const ActSrc& s, long long r, int c) {
    float v = ldf<T>(reinterpret_cast<const T*>(s.x1) + r * s.ld1 + c);
    if (s.a1) v *= s.a1[c];
    if (s.b1) v += s.b1[c];
    if (s.x2) {
        float w = ldf<T>(reinterpret_cast<const T*>(s.x2) + r * s.ld2 + c);
        if (s.a2) w *= s.a2[c];
        v += w;
    }
    if (s.b2) v += s.b2[c];
    if (s.relu) v = fmaxf(v, 0.f);
    return v;
}

// 4 consecutive channels c..c+3 (c % 4 == 0, ld % 4 == 0, base pointers 16B aligned)
template <class T> DSG_D float4 act_value4(const ActSrc& s, long long r, int c) {
    float4 v = ldf4<T>(reinterpret_cast<const T*>(s.x1) + r * s.ld1 + c);
    if (s.a1) { float4 a = *reinterpret_cast<const float4*>(s.a1 + c); v.x *= a.x; v.y *= a.y; v.z *= a.z; v.w *= a.w; }
    if (s.b1) { float4 b = *reinterpret_cast<const float4*>(s.b1 + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (s.x2) {
        float4 w = ldf4<T>(reinterpret_cast<const T*>(s.x2) + r * s.ld2 + c);
        if (s.a2) { float4 a = *reinterpret_cast<const float4*>(s.a2 + c); w.x *= a.x; w.y *= a.y; w.z *= a.z; w.w *= a.w; }
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (s.b2) { float4 b = *reinterpret_cast<const float4*>(s.b2 + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (s.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
}

static inline bool act_src_vec4_ok(const ActSrc& s, int elem_bytes) {
    size_t al = (size_t)(4 * elem_bytes);
    bool ok = ((uintptr_t)s.x1 % al == 0) && (s.ld1 % 4 == 0);
    if (s.x2) ok = ok && ((uintptr_t)s.x2 % al == 0) && (s.ld2 % 4 == 0);
    const float* cs[4] = {s.a1, s.b1, s.a2, s.b2};
    for (int i = 0; i < 4; ++i) if (cs[i]) ok = ok && ((uintptr_t)cs[i] % 16 == 0);
    return ok;
}

// ---- 8-channel (16-byte) vector helpers for bf16 activations -----------------------------------
DSG_D void unpack8(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
DSG_D uint4 pack8(const float* v) {
    uint4 u;
    u.x = dsg_pack_bf16x2(v[0], v[1]); u.y = dsg_pack_bf16x2(v[2], v[3]);
    u.z = dsg_pack_bf16x2(v[4], v[5]); u.w = dsg_pack_bf16x2(v[6], v[7]);
    return u;
}
DSG_D void load8f(const float* p, float* v, float dflt) {
    if (p) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = dflt;
    }
}
// register-resident coefficients of an activation source for one 8-channel chunk
struct Act8 {
    float a1[8], b[8], a2[8];     // b = b1 + b2
    const bf16* x1; const bf16* x2; long long ld1, ld2; int relu;
    DSG_D void init(const ActSrc& s, int c) {
        float t[8];
        load8f(s.a1 ? s.a1 + c : nullptr, a1, 1.f);
        load8f(s.b1 ? s.b1 + c : nullptr, b, 0.f);
        load8f(s.b2 ? s.b2 + c : nullptr, t, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] += t[j];
        load8f(s.a2 ? s.a2 + c : nullptr, a2, 1.f);
        x1 = reinterpret_cast<const bf16*>(s.x1) + c;
        x2 = s.x2 ? reinterpret_cast<const bf16*>(s.x2) + c : nullptr;
        ld1 = s.ld1; ld2 = s.ld2; relu = s.relu;
    }
    DSG_D void eval(long long r, float* v) const {
        float t[8];
        unpack8(*reinterpret_cast<const uint4*>(x1 + r * ld1), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(t[j], a1[j], b[j]);
        if (x2) {
            unpack8(*reinterpret_cast<const uint4*>(x2 + r * ld2), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(t[j], a2[j], v[j]);
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
    }
};

static inline bool act8_ok(const ActSrc& s) {
    bool ok = ((uintptr_t)s.x1 % 16 == 0) && (s.ld1 % 8 == 0);
    if (s.x2) ok = ok && ((uintptr_t)s.x2 % 16 == 0) && (s.ld2 % 8 == 0);
    const float* cs[4] = {s.a1, s.b1, s.a2, s.b2};
    for (int i = 0; i < 4; ++i) if (cs[i]) ok = ok && ((uintptr_t)cs[i] % 16 == 0);
    return ok;
}


