// scan3d_fused_kernel.cu -- the hot path: ONE persistent kernel that takes the captured pattern
// stack to absolute phase maps, fringe orders, c_p_map, validity and the raster-ordered compacted
// point cloud, reading every input byte once and writing every contract output once.
//
// Replaces, per scan, the reference call sequence compute_wrapped_phase(0/1) -> unwrap_phase(0/1)
// -> compute_c_p_map() -> triangulate() -> save_point_cloud()'s gather
// (M_tech_project_console/m_tech_project_console.cpp:372-401).
//
// Shape of the kernel (sm_100a):
//   * persistent grid, one CTA per SM, static round-robin over row-segment tiles of TILE_W pixels
//     (tile order == raster order, which the compaction relies on);
//   * warp-specialised: the last warp is the PRODUCER -- it moves each tile's NF frame segments
//     and a 4-row ROI window HBM -> shared memory with cp.async.bulk (TMA bulk copies) signalled
//     on an mbarrier ring of `stages` slots; the other 8 warps are CONSUMERS;
//   * consumers: (1) SWAR integer phase on 4 consecutive pixels per thread straight from shared
//     memory (mask recurrence closed form, phase-shift numerators/denominators, Gray threshold,
//     packed fringe orders); (2) FP64 phase per pixel with the reference's exact IEEE operation
//     order (atan2 -> +Pi -> +code*2Pi -> lrint correspondence -> undistorted pixels -> 4x3
//     normal-equation solve); (3) results are staged in the tile's own (now dead) shared-memory
//     slot and leave as TMA bulk stores; (4) valid points are compacted in raster order with a
//     block scan + decoupled look-back across tiles (epoch-tagged, no per-launch reset).
#include "scan3d_internal.h"

namespace s3d {

constexpr int TILE_W = 1024;               // pixels per tile (one row segment)
constexpr int NCONS = 256;                 // consumer threads, 4 pixels each
constexpr int NTHREADS = NCONS + 32;       // + one producer warp
constexpr int ROI_HALO = 16;               // bytes of halo each side (16 B aligned bulk copies)
constexpr int ROI_ROW = TILE_W + 2 * ROI_HALO;
constexpr int ROI_BYTES = 4 * ROI_ROW;     // rows y-2 .. y+1

// output staging offsets inside a stage slot (all 16 B aligned)
constexpr int OUT_UNWV = 0;                        // f32[1024]
constexpr int OUT_CODEV = OUT_UNWV + 4 * TILE_W;   // i16[1024]
constexpr int OUT_VALID = OUT_CODEV + 2 * TILE_W;  // u8[1024]
constexpr int OUT1_BYTES = OUT_VALID + TILE_W;     // dirs == 1 ends here
constexpr int OUT_UNWH = OUT1_BYTES;               // f32[1024]
constexpr int OUT_CODEH = OUT_UNWH + 4 * TILE_W;   // i16[1024]
constexpr int OUT_CP = OUT_CODEH + 2 * TILE_W;     // int2[1024]
constexpr int OUT_X = OUT_CP + 8 * TILE_W;         // f32[1024][3]
constexpr int OUT2_BYTES = OUT_X + 12 * TILE_W;

constexpr int SMEM_MAX = 227 * 1024;
constexpr int SMEM_FIXED = 8 * 16 + 66 * 8 + 64;   // barriers (<= 8 stages) + atan table + scratch

static int num_frames(const scan3d_config& c)
{
    return c.dirs == 2 ? 2 * c.N + 2 * (c.M_v + c.M_h) : c.N + 2 * c.M_v;
}
static int stage_bytes_of(const scan3d_config& c)
{
    const int in_b = num_frames(c) * TILE_W;
    const int out_b = c.dirs == 2 ? OUT2_BYTES : OUT1_BYTES;
    return (in_b > out_b ? in_b : out_b) + ROI_BYTES;
}

bool fused_supported(const scan3d_config& c, int* stages_out, size_t* smem_out)
{
    if (c.W % 16 != 0) return false;
    if (!(c.N == 3 || c.N == 4 || c.N == 5 || c.N == 8)) return false;
    const int sb = stage_bytes_of(c);
    int stages = (SMEM_MAX - SMEM_FIXED) / sb;
    if (stages > 4) stages = 4;
    if (stages < 2) return false;
    if (stages_out) *stages_out = stages;
    if (smem_out) *smem_out = (size_t)stages * sb + SMEM_FIXED;
    return true;
}

int fused_num_tiles(const scan3d_config& c)
{
    return ((c.W + TILE_W - 1) / TILE_W) * c.H;
}

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
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
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

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
// per-byte unsigned a >= b  ->  bit 7 of each byte (other bits garbage-free: masked)
__device__ __forceinline__ uint32_t ge_bytes(uint32_t a, uint32_t b)
{
    const uint32_t d = (a | 0x80808080u) - (b & 0x7f7f7f7fu);
    return ((a & ~b) | (~(a ^ b) & d)) & 0x80808080u;
}
// bytes (b0,b1,b2,b3) -> 16-bit lanes (b0,b1) and (b2,b3)
__device__ __forceinline__ uint32_t lanes_lo(uint32_t w) { return __byte_perm(w, 0, 0x4140); }
__device__ __forceinline__ uint32_t lanes_hi(uint32_t w) { return __byte_perm(w, 0, 0x4342); }

struct Terms {            // up to 4 biased 16-bit terms for 4 pixels: [term][0]=(px0,px1) [1]=(px2,px3)
    uint32_t t[4][2];
};
// phase-shift numerators/denominators for 4 pixels (3/wrapped_phase.cpp:171-173,195-196,217-218)
template <int N>
__device__ __forceinline__ void fringe_terms(const uint32_t* __restrict__ sw, int f0, int tid, Terms& T)
{
    uint32_t L[N][2];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint32_t w = sw[(f0 + k) * (TILE_W / 4) + tid];
        L[k][0] = lanes_lo(w);
        L[k][1] = lanes_hi(w);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (N == 3) {          // t1 = I0 - I2 (+512) ; t2 = 2*I1 - I0 - I2 (+1024)
            T.t[0][h] = L[0][h] + 0x02000200u - L[2][h];
            T.t[1][h] = 2 * L[1][h] + 0x04000400u - L[0][h] - L[2][h];
        } else if (N == 4) {   // t1 = I3 - I1 ; t2 = I0 - I2
            T.t[0][h] = L[3][h] + 0x02000200u - L[1][h];
            T.t[1][h] = L[0][h] + 0x02000200u - L[2][h];
        } else if (N == 5) {   // t1 = 2(I1 - I3) (+1024) ; t2 = 2*I2 - I0 - I4 (+1024)
            T.t[0][h] = 2 * L[1][h] + 0x04000400u - 2 * L[3][h];
            T.t[1][h] = 2 * L[2][h] + 0x04000400u - L[0][h] - L[4][h];
        } else {               // N == 8: a1 = I6-I2, b1 = I5+I7-I1-I3, a2 = I0-I4, b2 = I1+I7-I3-I5
            T.t[0][h] = L[6 % N][h] + 0x02000200u - L[2][h];
            T.t[1][h] = L[5 % N][h] + L[7 % N][h] + 0x04000400u - L[1][h] - L[3][h];
            T.t[2][h] = L[0][h] + 0x02000200u - L[4 % N][h];
            T.t[3][h] = L[1][h] + L[7 % N][h] + 0x04000400u - L[3][h] - L[5 % N][h];
        }
    }
}
__device__ __forceinline__ int term_of(const Terms& T, int k, int j, int bias)
{
    const uint32_t r = (j & 2) ? T.t[k][1] : T.t[k][0];
    return (int)((r >> ((j & 1) * 16)) & 0xffffu) - bias;
}

// Gray threshold for 4 pixels, all M planes: byte accumulators with plane i at bit (7 - i%8)
// (4/phase_unwrap.cpp:183: (uchar)img - (uchar)inv >= 0, tie -> 1)
__device__ __forceinline__ void gray_bits(const uint32_t* __restrict__ sw, int g0, int i0, int M, int tid,
                                          uint32_t& accA, uint32_t& accB)
{
    accA = 0; accB = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < M) accA |= ge_bytes(sw[(g0 + i) * (TILE_W / 4) + tid], sw[(i0 + i) * (TILE_W / 4) + tid]) >> i;
#pragma unroll
    for (int i = 8; i < 15; i++)
        if (i < M) accB |= ge_bytes(sw[(g0 + i) * (TILE_W / 4) + tid], sw[(i0 + i) * (TILE_W / 4) + tid]) >> (i - 8);
}
// Gray -> binary (B0 = G0, Bi = B(i-1) xor Gi) as a prefix xor; code = sum Bi << (M-1-i)  (:187-193)
__device__ __forceinline__ int code_of(uint32_t accA, uint32_t accB, int j, int M)
{
    const uint32_t a = (accA >> (8 * j)) & 0xffu, b = (accB >> (8 * j)) & 0xffu;
    uint32_t g = ((a << 8) | b) >> (16 - M);
    g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8;
    return (int)g;
}

template <int N>
__device__ __forceinline__ float phase_of(const Terms& T, int j, const double* tab)
{
    if (N == 8) {
        const int a1 = term_of(T, 0, j, 512), b1 = term_of(T, 1, j, 1024);
        const int a2 = term_of(T, 2, j, 512), b2 = term_of(T, 3, j, 1024);
        const double r = 0.70710678118654752440;
        const double d1 = dadd((double)a1, dmul((double)b1, r));
        const double d2 = dadd((double)a2, dmul((double)b2, r));
        const float f1 = fmaf((float)b1, 0.70710678f, (float)a1);
        const float f2 = fmaf((float)b2, 0.70710678f, (float)a2);
        return atan2_to_float(d1, d2, f1, f2, tab, tab + 33);
    } else {
        const int t1 = term_of(T, 0, j, N == 5 ? 1024 : 512);
        const int t2 = term_of(T, 1, j, N == 4 ? 512 : 1024);
        if (N == 5) return atan2f_fdlibm((float)t1, (float)t2);   // 3/wrapped_phase.cpp:220 (float atan2f)
        return atan2_to_float((double)t1, (double)t2, (float)t1, (float)t2, tab, tab + 33);
    }
}

// ---- the kernel --------------------------------------------------------------------------------
template <int N, int DIRS>
__global__ void __launch_bounds__(NTHREADS, 1)
k_fused(const __grid_constant__ FusedArgs a, const __grid_constant__ DeviceCalib cal, const int stages,
        const int stage_bytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    double* tab = reinterpret_cast<double*>(bars + 16);
    uint32_t* scr = reinterpret_cast<uint32_t*>(tab + 66);   // [0..7] warp sums, [8] tile base
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    const int W = a.W;
    const size_t plane = (size_t)W * a.H;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    for (int i = tid; i < 66; i += NTHREADS) tab[i] = a.atan_tab[i];
    __syncthreads();

    const int first = blockIdx.x;
    const int my_tiles = first < a.n_tiles ? (a.n_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == NCONS / 32) {
        // ================================ PRODUCER ================================
        for (int it = 0; it < my_tiles; it++) {
            const int tile = first + it * (int)gridDim.x;
            const int s = it % stages, ph = (it / stages) & 1;
            const int row = tile / a.tiles_per_row, seg = tile - row * a.tiles_per_row;
            const int x0 = seg * TILE_W, wt = min(TILE_W, W - x0);
            const int y = a.row0 + row;
            const int rx0 = max(0, x0 - ROI_HALO), rx1 = min(W, x0 + wt + ROI_HALO);
            const uint32_t rbytes = (uint32_t)(rx1 - rx0);
            int nrows = 0;
            for (int r = 0; r < 4; r++) nrows += (y - 2 + r >= 0 && y - 2 + r < a.H_total);
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
            if (lane == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)NF * wt + nrows * rbytes);
            __syncwarp();
            const uint8_t* src = a.stack + (size_t)row * W + x0;
            for (int f = lane; f < NF; f += 32)
                bulk_g2s(dst + f * TILE_W, src + (size_t)f * plane, (uint32_t)wt, bar_full + 8 * s);
            if (lane < 4) {
                const int yy = y - 2 + lane;
                if (yy >= 0 && yy < a.H_total)
                    bulk_g2s(dst + (stage_bytes - ROI_BYTES) + lane * ROI_ROW + (rx0 - (x0 - ROI_HALO)),
                             a.roi + (size_t)yy * W + rx0, rbytes, bar_full + 8 * s);
            }
        }
        return;
    }

    // ================================ CONSUMERS ================================
    const int store_tid = 32;   // warp 1 lane 0 issues the TMA stores and frees the slot
    int pending_release = -1;   // stage whose stores were issued by store_tid and not yet drained
    for (int it = 0; it < my_tiles; it++) {
        const int tile = first + it * (int)gridDim.x;
        const int s = it % stages, ph = (it / stages) & 1;
        const int row = tile / a.tiles_per_row, seg = tile - row * a.tiles_per_row;
        const int x0 = seg * TILE_W, wt = min(TILE_W, W - x0);
        const int y = a.row0 + row;
        const int xt = x0 + 4 * tid;            // first of this thread's 4 pixels (global column)
        const bool active = 4 * tid < wt;
        uint8_t* stage = smem + (size_t)s * stage_bytes;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage);
        const uint8_t* sroi = stage + (stage_bytes - ROI_BYTES);

        mbar_wait(bar_full + 8 * s, ph);

        // ---------------- integer phase: mask, fringe terms, Gray bits ----------------
        uint32_t mbits = 0;
        Terms Tv, Th;
        uint32_t gvA = 0, gvB = 0, ghA = 0, ghB = 0;
        if (active) {
            // ROI byte of global pixel (gx, gy), gy in [y-2, y+1], gx in [x0-16, x0+wt+16)
            auto inv = [&](int gx, int gy) { return sroi[(gy - (y - 2)) * ROI_ROW + (gx - x0 + ROI_HALO)] == 0; };
            const bool window_ok = y >= 2 && y + 1 < a.H_total && xt >= 4 && xt + 7 < W;
            bool fast = false;
            if (window_ok) {
                uint32_t any_zero = 0;   // haszero(v): non-zero iff some byte of v is zero
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(sroi + r * ROI_ROW + (4 * tid + ROI_HALO - 4) + 4 * c);
                        any_zero |= (v - 0x01010101u) & ~v & 0x80808080u;
                    }
                if (any_zero == 0) { mbits = 0xf; fast = true; }
            }
            if (!fast) {
                const uint32_t centre = *reinterpret_cast<const uint32_t*>(sroi + 2 * ROI_ROW + 4 * tid + ROI_HALO);
                if (centre != 0) {
#pragma unroll 1
                    for (int j = 0; j < 4; j++) {
                        const int x = xt + j;
                        bool v = !inv(x, y);
                        const bool border = x == 0 || y == 0 || x == W - 1 || y == a.H_total - 1;
                        if (v && !border) v = !mask_trigger(x, y, W, a.H_total, inv);
                        mbits |= (v ? 1u : 0u) << j;
                    }
                }
            }
            if (mbits) {
                fringe_terms<N>(sw, 0, tid, Tv);
                gray_bits(sw, N, N + a.M_v, a.M_v, tid, gvA, gvB);
                if (DIRS == 2) {
                    const int fh = N + 2 * a.M_v;
                    fringe_terms<N>(sw, fh, tid, Th);
                    gray_bits(sw, fh + N, fh + N + a.M_h, a.M_h, tid, ghA, ghB);
                }
            }
        }
        // barrier: every consumer is done reading this slot's inputs (they are overwritten by the
        // staged outputs below) and done with the previous tile's scatter, so store_tid can now
        // hand the previous slot back to the producer once its TMA stores have drained.
        cons_sync();
        if (tid == store_tid && pending_release >= 0) {
            bulk_wait_read();
            mbar_arrive(bar_empty + 8 * pending_release);
            pending_release = -1;
        }

        // ---------------- FP64 phase, one pixel at a time, results staged in the slot ----------------
        float* o_unwv = reinterpret_cast<float*>(stage + OUT_UNWV);
        int16_t* o_codev = reinterpret_cast<int16_t*>(stage + OUT_CODEV);
        uint8_t* o_valid = stage + OUT_VALID;
        float* o_unwh = reinterpret_cast<float*>(stage + OUT_UNWH);
        int16_t* o_codeh = reinterpret_cast<int16_t*>(stage + OUT_CODEH);
        int2* o_cp = reinterpret_cast<int2*>(stage + OUT_CP);
        float* o_x = reinterpret_cast<float*>(stage + OUT_X);
        uint32_t vbits = 0;
        if (active) {
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                const int x = xt + j, lp = 4 * tid + j;
                float unwv = 0.0f, unwh = 0.0f;
                int cv = -1, ch = -1;
                int2 cp = make_int2(0, 0);
                bool v = (mbits >> j) & 1u;
                if (v) {
                    cv = code_of(gvA, gvB, j, a.M_v);
                    if (!(x == 0 || x == W - 1))                                  // 4/phase_unwrap.cpp:285
                        unwv = unwrap_abs(add_pi(phase_of<N>(Tv, j, tab)), cv);   // :290-291
                    if (DIRS == 2) {
                        ch = code_of(ghA, ghB, j, a.M_h);
                        if (!(y == 0 || y == a.H_total - 1))                      // :304
                            unwh = unwrap_abs(add_pi(phase_of<N>(Th, j, tab)), ch);
                        long long px = 0, py = 0;                                 // 5/compute_correspondance.cpp:648-675
                        if (!correspond(unwv, a.fw_v, &px)) {
                            v = false;
                        } else if (!correspond(unwh, a.fw_h, &py)) {
                            v = false;
                            cp.x = (int)max(min(px, 2147483647LL), -2147483648LL);
                        } else {
                            cp.x = (int)max(min(px, 2147483647LL), -2147483648LL);
                            cp.y = (int)max(min(py, 2147483647LL), -2147483648LL);
                            if (px > a.PW - 1 || py > a.PH - 1 || px < 0 || py < 0) v = false;
                        }
                    }
                }
                o_unwv[lp] = unwv;
                o_codev[lp] = (int16_t)cv;
                if (DIRS == 2) {
                    o_unwh[lp] = unwh;
                    o_codeh[lp] = (int16_t)ch;
                    o_cp[lp] = cp;
                    if (v) {                                                      // 7/triangulation.cpp:1230-1247
                        double uc, vc, up, vp, X[3];
                        if (a.cam_lut) {
                            const double2 t = a.cam_lut[(size_t)row * W + x];
                            uc = t.x; vc = t.y;
                        } else {
                            undistorted_pixel_nodist(cal.Kc, (double)x, (double)y, &uc, &vc);
                        }
                        if (a.proj_lut) {
                            const double2 t = a.proj_lut[(size_t)cp.y * a.PW + cp.x];
                            up = t.x; vp = t.y;
                        } else {
                            undistorted_pixel_nodist(cal.Kp, (double)cp.x, (double)cp.y, &up, &vp);
                        }
                        triangulate_point(cal.Ac, cal.Ap, uc, vc, up, vp, X);
                        o_x[3 * lp + 0] = __double2float_rn(X[0]);               // 8/save_point_cloud.cpp:94-96
                        o_x[3 * lp + 1] = __double2float_rn(X[1]);
                        o_x[3 * lp + 2] = __double2float_rn(X[2]);
                    }
                }
                o_valid[lp] = v ? 1 : 0;
                vbits |= (v ? 1u : 0u) << j;
            }
        }
        fence_async_smem();   // staged outputs -> visible to the async (TMA) proxy

        // ---------------- block scan of valid counts ----------------
        uint32_t excl_in_tile = 0;
        if (DIRS == 2) {
            const uint32_t cnt = __popc(vbits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) scr[warp] = incl;
            excl_in_tile = incl - cnt;
        }
        cons_sync();

        // ---------------- TMA stores of the plane outputs (one thread) ----------------
        if (tid == store_tid) {
            const size_t g = (size_t)row * W + x0;
            bulk_s2g(a.unw_v + g, smem_u32(stage + OUT_UNWV), 4u * wt);
            bulk_s2g(a.code_v + g, smem_u32(stage + OUT_CODEV), 2u * wt);
            bulk_s2g(a.valid + g, smem_u32(stage + OUT_VALID), 1u * wt);
            if (DIRS == 2) {
                bulk_s2g(a.unw_h + g, smem_u32(stage + OUT_UNWH), 4u * wt);
                bulk_s2g(a.code_h + g, smem_u32(stage + OUT_CODEH), 2u * wt);
                bulk_s2g(a.cpmap + g, smem_u32(stage + OUT_CP), 8u * wt);
            }
            bulk_commit();
            pending_release = s;
        }

        if (DIRS == 2) {
            // ---------------- decoupled look-back across tiles (warp 0) ----------------
            if (warp == 0) {
                uint32_t wsum = lane < NCONS / 32 ? scr[lane] : 0;
                uint32_t total = wsum;
#pragma unroll
                for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                const unsigned long long tag = (unsigned long long)(a.epoch & 0x3fffffffu) << 34;
                uint32_t excl = 0;
                if (tile > 0) {
                    if (lane == 0) st_state(a.tile_state + tile, tag | (1ull << 32) | total);
                    int look = tile - 1;
                    while (true) {
                        const int idx = look - lane;
                        unsigned long long w = tag | (2ull << 32);   // virtual tile < 0: prefix 0
                        if (idx >= 0) {
                            do {
                                w = ld_state(a.tile_state + idx);
                            } while ((w >> 34) != (tag >> 34) || ((w >> 32) & 3ull) == 0);
                        }
                        const bool is_prefix = ((w >> 32) & 3ull) == 2;
                        const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                        const int stop = pm ? __ffs(pm) - 1 : 31;     // nearest tile holding a prefix
                        uint32_t v = lane <= stop ? (uint32_t)w : 0;
#pragma unroll
                        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        excl += v;
                        if (pm) break;
                        look -= 32;
                    }
                }
                if (lane == 0) {
                    st_state(a.tile_state + tile, tag | (2ull << 32) | (excl + total));
                    scr[8] = excl;
                    if (tile == a.n_tiles - 1) *a.d_count = excl + total;
                }
            }
            cons_sync();

            // ---------------- raster-ordered scatter of this tile's points ----------------
            if (vbits) {
                uint32_t warp_off = 0;
                for (int w2 = 0; w2 < warp; w2++) warp_off += scr[w2];
                uint32_t dst = scr[8] + warp_off + excl_in_tile;
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    if (!((vbits >> j) & 1u)) continue;
                    const int lp = 4 * tid + j;
                    float* o = a.pts + 3 * (size_t)dst;
                    o[0] = o_x[3 * lp + 0];
                    o[1] = o_x[3 * lp + 1];
                    o[2] = o_x[3 * lp + 2];
                    const size_t gp = (size_t)row * W + x0 + lp;
                    if (a.pix) a.pix[dst] = (uint32_t)((size_t)a.row0 * W + gp);
                    if (a.rgb) {
                        a.rgb[3 * (size_t)dst + 0] = a.texture[3 * gp + 2];
                        a.rgb[3 * (size_t)dst + 1] = a.texture[3 * gp + 1];
                        a.rgb[3 * (size_t)dst + 2] = a.texture[3 * gp + 0];
                    }
                    dst++;
                }
            }
            // the slot is released by store_tid after the next tile's first barrier, which orders
            // these shared-memory reads before the producer's refill.
        }
    }
    if (tid == store_tid) {
        if (pending_release >= 0) {
            bulk_wait_read();
            mbar_arrive(bar_empty + 8 * pending_release);
        }
        bulk_wait_all();
    }
}

template <int N, int DIRS>
static cudaError_t launch_fused_t(const FusedArgs& a, const DeviceCalib& cal, int sm_count, int stages,
                                  int stage_bytes, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_fused<N, DIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    k_fused<N, DIRS><<<grid, NTHREADS, smem, st>>>(a, cal, stages, stage_bytes);
    return cudaGetLastError();
}

cudaError_t launch_fused(const scan3d_config& c, const FusedArgs& a_in, const DeviceCalib& cal, int sm_count,
                         cudaStream_t st)
{
    int stages = 0;
    size_t smem = 0;
    if (!fused_supported(c, &stages, &smem)) return cudaErrorInvalidValue;
    if (((uintptr_t)a_in.stack & 15) || ((uintptr_t)a_in.roi & 15)) return cudaErrorMisalignedAddress;
    FusedArgs a = a_in;
    a.tiles_per_row = (c.W + TILE_W - 1) / TILE_W;
    a.n_tiles = a.tiles_per_row * c.H;
    const int sb = stage_bytes_of(c);
#define S3D_F(NN)                                                                              \
    (c.dirs == 2 ? launch_fused_t<NN, 2>(a, cal, sm_count, stages, sb, smem, st)               \
                 : launch_fused_t<NN, 1>(a, cal, sm_count, stages, sb, smem, st))
    switch (c.N) {
        case 3: return S3D_F(3);
        case 4: return S3D_F(4);
        case 5: return S3D_F(5);
        case 8: return S3D_F(8);
    }
#undef S3D_F
    return cudaErrorInvalidValue;
}

}  // namespace s3d
