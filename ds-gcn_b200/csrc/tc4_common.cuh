// Shared vocabulary of the TMA-fed tcgen05 engines (tc4_gemm.cuh, tc4_wgrad.cuh) — sm_100a only.
//
//  * operand tiles are moved by the TMA unit (cp.async.bulk.tensor.{2d,3d}, SWIZZLE_128B) straight into the layout the
//    tensor core reads: "atoms" of [rows x 64 bf16 channels] = rows x 128 B, 16-byte chunk index XOR-ed with (row & 7),
//    every atom 1024-byte aligned.  The same atom is a K-major operand (reduction over channels: start + 32 B per K = 16,
//    SBO = 1024) and an MN-major operand (reduction over rows: start + 2048 B per 16 rows, SBO = 1024, LBO = distance
//    between 64-channel atoms) — verified on the B200 by tools/probes/probe_tc4.cu, including row-shifted start addresses;
//  * out-of-bounds box elements (channel tails, rows past the end, the padding joint of a 3-D [C, V, frames] view) are
//    zero-filled on load and clipped on store by the TMA unit, so no kernel carries boundary code;
//  * tensor maps are encoded on the host per call (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: the library
//    does not link libcuda) and passed as __grid_constant__ kernel parameters, so they are captured by value in CUDA graphs.
#pragma once
#include "conv_gemm_tc3.cuh"

#ifndef DSG_EMU
#include <cuda.h>
namespace dsg {
namespace tc4 {

using tc::smem_u32;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::bulk_g2s;
using tc::umma_f16;
using tc::umma_commit;
using tc::tmem_alloc;
using tc::tmem_dealloc;
using tc::tmem_ld16;
using tc::make_idesc;

constexpr int ATOM_CH = 64;                      // channels per atom (128 bytes of bf16)
constexpr int ATOM_ROWS = 128;
constexpr int ATOM_BYTES = ATOM_ROWS * 128;      // 16 KB

// ---- host: tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
static inline bool tma_ptr_ok(const void* p, long long ld) { return p != nullptr && (uintptr_t)p % 16 == 0 && ld % 8 == 0 && ld > 0; }

// bf16 activation [rows, C] (pitch ld) seen as 2-D (C, rows); box = (64, box_rows)
static inline bool make_map_2d(CUtensorMap* m, const void* base, long long rows, int C, long long ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || rows <= 0 || C <= 0) return false;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)ATOM_CH, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the same buffer seen as 3-D (C, rows_per_frame, frames); box = (64, box_v, box_f): a box is dense [box_f][box_v] rows in
// shared memory, and box_v > rows_per_frame zero-fills (load) / skips (store) the padding rows
static inline bool make_map_3d(CUtensorMap* m, const void* base, long long frames, int rpf, int C, long long ld, int box_v, int box_f) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || frames <= 0 || C <= 0 || rpf <= 0) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rpf, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rpf};
    cuuint32_t box[3] = {(cuuint32_t)ATOM_CH, (cuuint32_t)box_v, (cuuint32_t)box_f};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D view (C, rows_per_frame, frames of one temporal plane, samples) of [n, T, rpf, C]: frame pitch `t_stride` frames (temporal
// stride / parity planes), first frame `t0`; box = (64, rpf, 1, 1).  A frame coordinate outside [0, frames) is zero-filled on
// load — exactly the zero padding of a temporal convolution — and skipped on store.
static inline bool make_map_4d(CUtensorMap* m, const void* base, int n_samples, int T, int rpf, int C, long long ld, int t0, int t_stride,
                               int box_f = 1) {
    EncodeTiledFn enc = encode_fn();
    const int frames = T > t0 ? (T - t0 + t_stride - 1) / t_stride : 0;
    if (!enc || n_samples <= 0 || frames <= 0 || C <= 0 || rpf <= 0) return false;
    const unsigned char* b0 = reinterpret_cast<const unsigned char*>(base) + (size_t)t0 * rpf * ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)rpf, (cuuint64_t)frames, (cuuint64_t)n_samples};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rpf * (cuuint64_t)t_stride, (cuuint64_t)ld * 2 * (cuuint64_t)rpf * (cuuint64_t)T};
    cuuint32_t box[4] = {(cuuint32_t)ATOM_CH, (cuuint32_t)rpf, (cuuint32_t)box_f, 1};       // box_f frames land densely packed: [box_f][rpf] rows
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<unsigned char*>(b0), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device: TMA --------------------------------------------------------------------------------------------------
DSG_D void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
DSG_D void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
DSG_D void tma_load_4d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
DSG_D void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
DSG_D void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
DSG_D void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
DSG_D void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> DSG_D void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> DSG_D void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
DSG_D void prefetch_map(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }
DSG_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DSG_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DSG_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DSG_D void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Waiting with back-off: in the fused-contraction mode four transform warps do long CUDA-core work while ten other warps of the CTA
// wait on mbarriers; a bare try_wait loop re-issues every ~10 ns and (ncu, profiles/r02_ncu_full_tc4_fused_contraction.json) took two
// thirds of the SM's issue slots away from the working warps.  Sleeping between polls gives the slots back.
DSG_D void mbar_wait_backoff(uint64_t* bar, uint32_t parity, bool backoff) {
    if (!backoff) { mbar_wait(bar, parity); return; }
    uint32_t spins = 0;
    while (!tc::mbar_try_wait(bar, parity)) {
        __nanosleep(256);
        if (++spins > (1u << 24)) __trap();
    }
}

// ---- device: UMMA descriptors over SWIZZLE_128B atoms ---------------------------------------------------------------
// K-major (reduction over the 64 channels of the atom): rows 128 B apart, 8-row groups 1024 B apart
DSG_D uint64_t desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major (reduction over the rows of the tile): LBO = distance between 64-channel atoms, SBO = 1024 (8-row groups)
DSG_D uint64_t desc_mn_sw128(uint32_t saddr, uint32_t atom_stride) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((atom_stride >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// the "ones" operand: [8 x 16] bf16 ones, K-major no-swizzle (any layout reads ones)
DSG_D uint64_t desc_ones(uint32_t saddr) { return tc::make_desc(saddr, 128u, 256u); }
// instruction descriptor with explicit operand majors (bit 15: A is MN-major, bit 16: B is MN-major)
DSG_D uint32_t idesc_major(int M, int N, int a_mn, int b_mn) { return make_idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16); }

// byte offset of (row, 16-byte chunk) inside an atom
DSG_D uint32_t atom_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// TMEM lane that holds accumulator row m of an M = 64 (cta_group::1) MMA: 16 rows per 32-lane quarter (probe_tc4)
DSG_HD int m64_lane(int m) { return (m >> 4) * 32 + (m & 15); }

DSG_D void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

static inline bool tc4_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSG_DISABLE_TC4"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static inline int num_sms() { return dsg_num_sms(); }

}  // namespace tc4
}  // namespace dsg
#endif
