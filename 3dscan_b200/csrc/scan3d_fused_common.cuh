// scan3d_fused_common.cuh -- device helpers shared by the fused kernels: PTX wrappers (mbarrier,
// TMA bulk copies, look-back words), the SWAR integer decode of 4 pixels per 32-bit word and the
// pipeline timeline hooks.
#pragma once
#include "scan3d_internal.h"

namespace s3d {

constexpr int ROI_HALO = 16;               // bytes of halo each side (16 B aligned bulk copies)
constexpr int SMEM_MAX = 227 * 1024;

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// (A non-blocking mbarrier.test_wait probe in the IO warp's event loop measured 4-10 % SLOWER than
// try_wait, whose hardware suspend keeps the polling warp out of the issue slots.)
// poll with back-off so that a waiting warp does not steal issue slots from the computing ones
// (one copy per warp role so that profiles attribute the waiting to the right role)
#define S3D_WAIT_BODY                                        \
    if (mbar_try(bar, parity)) return;                      \
    while (!mbar_try(bar, parity)) __nanosleep(128);
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    S3D_WAIT_BODY
}
__device__ __forceinline__ void mbar_wait_consumer(uint32_t bar, uint32_t parity)
{
    S3D_WAIT_BODY
}
__device__ __forceinline__ void mbar_wait_epilogue(uint32_t bar, uint32_t parity)
{
    S3D_WAIT_BODY
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// one 2-D tiled tensor copy global -> shared through a CUtensorMap (coordinates in elements)
__device__ __forceinline__ void tensor_g2s_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int NT>
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- SWAR pieces -----------------------------------------------------------------------------
// per-byte unsigned a >= b  ->  bit 7 of each byte (the other bits are NOT cleared: callers mask)
__device__ __forceinline__ uint32_t ge_bytes_raw(uint32_t a, uint32_t b)
{
    const uint32_t d = (a | 0x80808080u) - (b & 0x7f7f7f7fu);
    return (a & ~b) | (~(a ^ b) & d);
}
__device__ __forceinline__ uint32_t ge_bytes(uint32_t a, uint32_t b) { return ge_bytes_raw(a, b) & 0x80808080u; }
// bytes (b0,b1,b2,b3) -> 16-bit lanes (b0,b1) and (b2,b3)
__device__ __forceinline__ uint32_t lanes_lo(uint32_t w) { return __byte_perm(w, 0, 0x4140); }
__device__ __forceinline__ uint32_t lanes_hi(uint32_t w) { return __byte_perm(w, 0, 0x4342); }

struct Terms {            // up to 4 biased 16-bit terms for 4 pixels: [term][0]=(px0,px1) [1]=(px2,px3)
    uint32_t t[4][2];
};
// phase-shift numerators/denominators for 4 pixels (3/wrapped_phase.cpp:171-173,195-196,217-218)
template <int N>
__device__ __forceinline__ void fringe_terms(const uint32_t* __restrict__ sw, int f0, int wpf, int tid, Terms& T)
{
    uint32_t L[N][2];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint32_t w = sw[(f0 + k) * wpf + tid];
        L[k][0] = lanes_lo(w);
        L[k][1] = lanes_hi(w);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (N == 3) {          // t1 = I0 - I2 (+512) ; t2 = 2*I1 - I0 - I2 (+1024)
            T.t[0][h] = L[0][h] + 0x02000200u - L[2][h];
            T.t[1][h] = 2 * L[1][h] + 0x04000400u - L[0][h] - L[2][h];
        } else if (N == 4) {   // t1 = I3 - I1 ; t2 = I0 - I2
            T.t[0][h] = L[3 % N][h] + 0x02000200u - L[1][h];
            T.t[1][h] = L[0][h] + 0x02000200u - L[2][h];
        } else if (N == 5) {   // t1 = 2(I1 - I3) (+1024) ; t2 = 2*I2 - I0 - I4 (+1024)
            T.t[0][h] = 2 * L[1][h] + 0x04000400u - 2 * L[3 % N][h];
            T.t[1][h] = 2 * L[2][h] + 0x04000400u - L[0][h] - L[4 % N][h];
        } else {               // N == 8: a1 = I6-I2, b1 = I5+I7-I1-I3, a2 = I0-I4, b2 = I1+I7-I3-I5
            T.t[0][h] = L[6 % N][h] + 0x02000200u - L[2][h];
            T.t[1][h] = L[5 % N][h] + L[7 % N][h] + 0x04000400u - L[1][h] - L[3 % N][h];
            T.t[2][h] = L[0][h] + 0x02000200u - L[4 % N][h];
            T.t[3][h] = L[1][h] + L[7 % N][h] + 0x04000400u - L[3 % N][h] - L[5 % N][h];
        }
    }
}
__device__ __forceinline__ int term_of(const Terms& T, int k, int j, int bias)
{
    const uint32_t r = (j & 2) ? T.t[k][1] : T.t[k][0];
    return (int)((r >> ((j & 1) * 16)) & 0xffffu) - bias;
}

// Gray -> binary (B0 = G0, Bi = B(i-1) xor Gi, 4/phase_unwrap.cpp:187-191) for the 4 pixels at
// once: a prefix xor inside every byte (plane i sits at bit 7 - i%8), then the parity of planes
// 0..7 (bit 0 of the first accumulator) carried into every bit of the second one.
__device__ __forceinline__ void gray_to_binary(uint32_t& accA, uint32_t& accB)
{
    accA ^= (accA >> 1) & 0x7f7f7f7fu;
    accA ^= (accA >> 2) & 0x3f3f3f3fu;
    accA ^= (accA >> 4) & 0x0f0f0f0fu;
    accB ^= (accB >> 1) & 0x7f7f7f7fu;
    accB ^= (accB >> 2) & 0x3f3f3f3fu;
    accB ^= (accB >> 4) & 0x0f0f0f0fu;
    accB ^= (accA & 0x01010101u) * 0xffu;
}
// Gray threshold for 4 pixels, all M planes: byte accumulators with plane i at bit (7 - i%8)
// (4/phase_unwrap.cpp:183: (uchar)img - (uchar)inv >= 0, tie -> 1)
__device__ __forceinline__ void gray_bits(const uint32_t* __restrict__ sw, int g0, int i0, int M, int wpf, int tid,
                                          uint32_t& accA, uint32_t& accB)
{
    accA = 0; accB = 0;
    // shift first, then mask and merge in one 3-input logic op: acc | ((ge >> i) & (0x80808080 >> i))
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < M) accA |= (ge_bytes_raw(sw[(g0 + i) * wpf + tid], sw[(i0 + i) * wpf + tid]) >> i) & (0x80808080u >> i);
#pragma unroll
    for (int i = 8; i < 15; i++)
        if (i < M) accB |= (ge_bytes_raw(sw[(g0 + i) * wpf + tid], sw[(i0 + i) * wpf + tid]) >> (i - 8)) & (0x80808080u >> (i - 8));
    gray_to_binary(accA, accB);
}
// fringe order of pixel j from the binary accumulators: code = sum Bi << (M-1-i)  (:193)
__device__ __forceinline__ int code_of(uint32_t binA, uint32_t binB, int j, int M)
{
    const uint32_t a = (binA >> (8 * j)) & 0xffu, b = (binB >> (8 * j)) & 0xffu;
    return (int)(((a << 8) | b) >> (16 - M));
}

// biased 16-bit lane -> exact double without the conversion pipe: the lane value u (< 2^16) is
// dropped into the mantissa of 2^52 and (2^52 + bias) is subtracted
__device__ __forceinline__ double term_f64(const Terms& T, int k, int j, int bias)
{
    const uint32_t r = (j & 2) ? T.t[k][1] : T.t[k][0];
    const uint32_t u = (j & 1) ? (r >> 16) : (r & 0xffffu);
    return __hiloint2double(0x43300000, (int)u) - (4503599627370496.0 + (double)bias);
}

template <int N>
__device__ __forceinline__ float phase_of(const Terms& T, int j, const double* tab)
{
    if (N == 8) {
        const double a1 = term_f64(T, 0, j, 512), b1 = term_f64(T, 1, j, 1024);
        const double a2 = term_f64(T, 2, j, 512), b2 = term_f64(T, 3, j, 1024);
        const double r = 0.70710678118654752440;
        const double d1 = dadd(a1, dmul(b1, r));
        const double d2 = dadd(a2, dmul(b2, r));
        // float approximations only pick the atan table row
        return atan2_to_float(d1, d2, __double2float_rz(d1), __double2float_rz(d2), tab);
    } else {
        const int t1 = term_of(T, 0, j, N == 5 ? 1024 : 512);
        const int t2 = term_of(T, 1, j, N == 4 ? 512 : 1024);
        if (N == 5) return atan2f_fdlibm((float)t1, (float)t2);   // 3/wrapped_phase.cpp:220 (float atan2f)
        return atan2_to_float((double)t1, (double)t2, (float)t1, (float)t2, tab);
    }
}

__device__ __forceinline__ int sat32(long long v)
{
    return (int)max(min(v, 2147483647LL), -2147483648LL);
}

// optional per-CTA timeline (library built with SCAN3D_BUILD_TRACE=1, run with SCAN3D_TRACE=1):
// clock64 stamps of the pipeline events of the first TRACE_TILES tiles of every CTA;
// slot = (cta * TRACE_TILES + it) * 8 + event
#ifndef S3D_TRACE
#define S3D_TRACE 0
#endif
constexpr int TRACE_TILES = 64;
__device__ __forceinline__ void trace(unsigned long long* t, int it, int ev)
{
    if (S3D_TRACE && t && it < TRACE_TILES) t[((size_t)blockIdx.x * TRACE_TILES + it) * 8 + ev] = clock64();
}


// work-list kernels (scan3d_fused_kernel.cu)
cudaError_t launch_worklist(const FusedArgs& a, int T, int dirs, cudaStream_t st);

}  // namespace s3d
