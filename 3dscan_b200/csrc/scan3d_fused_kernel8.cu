// scan3d_fused_kernel8.cu -- third cut of the single-pass kernel ("v8"): WARP-AUTONOMOUS pipelines.  Same work and
// the same bit-for-bit results as k_fused7 (scan3d_fused_kernel7.cu); what changed is who waits for whom.  The
// round-2 per-line profile of v7 (profiles/r2_v7_ncu_by_line.txt) shows the consumer warps 17 % of their time at
// the CTA-wide count barrier, 12 % on global loads and ~5 % waiting for the CTA's single input slot, at 62 % issue
// utilisation; two intermediate cuts that kept an IO warp per CTA (one slot per CTA, then one sub-slot per warp)
// only moved that waiting to the next shared resource (profiles/r2_optimisation_log.md).  So here NOTHING is shared
// between the warps of a CTA:
//
//   * the work unit is a CHUNK of 128 consecutive pixels of the row-major frame = one warp x 4 pixels per thread;
//     a warp draws chunk indices from a global counter that is never reset (the host passes the counter's value
//     at launch; every warp overdraws exactly once, so a launch advances it by n_chunks + warps) -- ONE launch per
//     scan, no work-list pre-pass;
//   * the drawing warp looks at the chunk's own ROI bytes: no selected pixel -> it writes the chunk's constant
//     outputs, publishes a zero count and draws again; the 56 frame segments of such a chunk are never read;
//   * every warp owns a sub-slot [frames] x [128 B] + a 4-row ROI window in shared memory and ONE mbarrier.  As
//     soon as its integer phase has consumed the sub-slot it issues the tensor-map copy of its NEXT chunk itself
//     (the draw for it was made a whole chunk earlier), which lands while the FP64 phase of the current one runs;
//   * raster-order compaction without anybody waiting: the warp publishes the chunk's survivor count right after
//     the decode, triangulates straight into a small ring of staging slots in global memory (L2 resident: the ring
//     is rewritten every few chunks) and only D chunks LATER resolves the chunk's place in the raster order --
//     decoupled look-back over the chunk counts, which by then are all there and mostly already prefixes -- and
//     moves the points to their final place.  Zero-count chunks go through the same deferred resolution, so runs of
//     empty chunks never lengthen anybody's look-back;
//   * no IO warp: 8 consumer warps per CTA, 3 CTAs = 24 warps per SM at 80 registers (v7: 21 + 3 IO warps).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scan3d_fused_common.cuh"

namespace s3d {

#ifndef S3D_K8_WARPS
#define S3D_K8_WARPS 8
#endif
#ifndef S3D_K8_MINB
#define S3D_K8_MINB 3
#endif
constexpr int K8_WARPS = S3D_K8_WARPS;           // warps per CTA, all of them workers
constexpr int K8_ROI_ROW = 128 + 2 * ROI_HALO;   // one row of a warp's ROI window: its 128 pixels + halo
constexpr int K8_RING = 32;                      // pending (published, not yet resolved) chunks per warp
constexpr int K8_D = 3;                          // a chunk's place in the raster order is resolved D chunks later
constexpr int K8_DS = K8_D + 1;                  // staging slots per warp

static int num_frames8(const scan3d_config& c)
{
    return c.dirs == 2 ? 2 * c.N + 2 * (c.M_v + c.M_h) : c.N + 2 * c.M_v;
}
static size_t smem8(const scan3d_config& c)
{
    return (size_t)K8_WARPS * ((size_t)num_frames8(c) * 128 + 4 * K8_ROI_ROW + K8_RING * 8 + 16) + ATAN_TAB_DOUBLES * 8;
}

struct Plan8 {
    int minb;
    size_t smem;
};

static bool plan8(const scan3d_config& c, Plan8* out)
{
    if (c.W % 16 != 0) return false;
    if (!(c.N == 3 || c.N == 4 || c.N == 5 || c.N == 8)) return false;
    int minb0 = S3D_K8_MINB;
    if (const char* e = getenv("SCAN3D_FUSED_CFG")) {
        const int b = atoi(e);
        if (b == 2 || b == S3D_K8_MINB) minb0 = b;
    }
    const size_t sm = smem8(c);
    for (int b = minb0; b >= 2; b -= (b == S3D_K8_MINB ? S3D_K8_MINB - 2 : 1))
        if ((sm + 1024) * b <= (size_t)SMEM_MAX + 1024) {
            out->minb = b; out->smem = sm;
            return true;
        }
    return false;
}

bool fused8_supported(const scan3d_config& c)
{
    Plan8 p;
    return plan8(c, &p);
}

int fused8_num_chunks(const scan3d_config& c) { return (int)(((size_t)c.W * c.H + 127) / 128); }

constexpr int regs8(int minb)
{
    // per SM sub-partition: resident warps / 4 share 16384 registers
    const int r = 16384 / (((minb * K8_WARPS + 3) / 4) * 32);
    return r > 255 ? 255 : (r / 8) * 8;
}

// mask recurrence for the 4 pixels of a thread next to a ROI edge or the frame border (rare): out of line
static __device__ __noinline__ uint32_t mask_slow8(const uint8_t* sroi, int lpw, int xt, int y, int W, int H_total)
{
    uint32_t mbits = 0;
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        const int x = xt + j;
        auto inv = [&](int gx, int gy) { return sroi[(gy - y + 2) * K8_ROI_ROW + (lpw + j + (gx - x) + ROI_HALO)] == 0; };
        bool v = !inv(x, y);
        const bool border = x == 0 || y == 0 || x == W - 1 || y == H_total - 1;
        if (v && !border) v = !mask_trigger(x, y, W, H_total, inv);
        mbits |= (v ? 1u : 0u) << j;
    }
    return mbits;
}

// ---- a pending chunk (ring entry e = {chunk, survivors | staging slot << 16 | age stamp << 24}) gets its place in
//      the raster order: decoupled look-back over the chunk states (all lower chunks' counts are normally long out,
//      most of them already as prefixes), the inclusive prefix is published, the chunk's points move from the warp's
//      staging slot to their final place.  Out of line: the hot loop stays small, and this runs once per chunk.
static __device__ __noinline__ void resolve8(const FusedArgs& a, int2 e, unsigned long long tag, const float* stage,
                                             const uint32_t* stage_vb, int n_chunks, int lane)
{
    const int c = e.x;
    const uint32_t n = (uint32_t)e.y & 0xffffu;
    const int slot = (e.y >> 16) & 0xff;
    uint32_t excl = 0;
    int look = c - 1;
    for (;;) {
        const int idx = look - lane;
        unsigned long long ws = tag | (2ull << 32);   // virtual chunk < 0: prefix 0
        if (idx >= 0) ws = ld_state(a.tile_state + idx);
        const bool okw = (ws >> 34) == (tag >> 34) && ((ws >> 32) & 3ull) != 0;
        const bool is_prefix = okw && ((ws >> 32) & 3ull) == 2;
        const unsigned rm = __ballot_sync(0xffffffffu, okw);
        const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
        // needed lanes: from the nearest chunk up to the first known prefix
        const int stop = pm ? __ffs(pm) - 1 : 31;
        const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
        if ((rm & need) != need) { __nanosleep(200); continue; }   // a lower chunk is still being decoded somewhere
        uint32_t v = lane <= stop ? (uint32_t)ws : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) {
        st_state(a.tile_state + c, tag | (2ull << 32) | (excl + n));
        if (c == n_chunks - 1) *a.d_count = excl + n;
    }
    if (n == 0) return;
    const float* src = stage + slot * 384;
    float* dst = a.pts + 3 * (size_t)excl;
    const int nf = 3 * (int)n;
    float v[12];
#pragma unroll
    for (int q = 0; q < 12; q++) {
        const int i = lane + 32 * q;
        v[q] = i < nf ? __ldcg(src + i) : 0.0f;
    }
#pragma unroll
    for (int q = 0; q < 12; q++) {
        const int i = lane + 32 * q;
        if (i < nf) dst[i] = v[q];
    }
    if (stage_vb) {
        const uint32_t* vb4 = stage_vb + slot * 4;
        const uint32_t vb = ((__ldcg(vb4 + 0) >> lane) & 1u) | (((__ldcg(vb4 + 1) >> lane) & 1u) << 1) |
                            (((__ldcg(vb4 + 2) >> lane) & 1u) << 2) | (((__ldcg(vb4 + 3) >> lane) & 1u) << 3);
        const uint32_t cn = __popc(vb);
        uint32_t incl = cn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t dp = excl + incl - cn;
        for (int j = 0; j < 4; j++)
            if ((vb >> j) & 1u) {
                const size_t gp = (size_t)c * 128 + 4 * lane + j;
                if (a.pix) a.pix[dp] = (uint32_t)((size_t)a.row0 * a.W + gp);
                if (a.rgb) {
                    a.rgb[3 * (size_t)dp + 0] = a.texture[3 * gp + 2];
                    a.rgb[3 * (size_t)dp + 1] = a.texture[3 * gp + 1];
                    a.rgb[3 * (size_t)dp + 2] = a.texture[3 * gp + 0];
                }
                dp++;
            }
    }
}

template <int N, int DIRS, int MINB, bool EXACT>
__global__ void __launch_bounds__(K8_WARPS * 32) __maxnreg__(regs8(MINB))
k_fused8(const __grid_constant__ FusedArgs a, const __grid_constant__ DeviceCalib cal, const __grid_constant__ CUtensorMap stack_map)
{
    constexpr bool fastdiv = true;   // host verified (else this kernel is not used)

    extern __shared__ __align__(128) uint8_t smem[];
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = a.W;
    const int plane = W * a.H;
    const int n_chunks = a.n_tiles;

    // shared memory: [warp] sub-slots | [warp] ROI windows | [warp] pending rings | [warp] mbarrier | atan table
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(smem + (size_t)warp * NF * 128);
    uint8_t* p_roi = smem + (size_t)K8_WARPS * NF * 128;
    const uint8_t* sroi = p_roi + warp * 4 * K8_ROI_ROW;
    int2* ring = reinterpret_cast<int2*>(p_roi + K8_WARPS * 4 * K8_ROI_ROW) + warp * K8_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_roi + K8_WARPS * (4 * K8_ROI_ROW + K8_RING * 8));
    double* tab = reinterpret_cast<double*>(bars + 2 * K8_WARPS);
    const uint32_t my_full = smem_u32(bars + warp);

    if (lane == 0) {
        mbar_init(my_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    for (int i = tid; i < ATAN_TAB_DOUBLES; i += K8_WARPS * 32) tab[i] = a.atan_tab[i];
    __syncthreads();
    // ---- from here on the warps of the CTA never meet again ----

    const unsigned long long tag = (unsigned long long)(a.epoch & 0x3fffffffu) << 34;
    const long long roi_total = (long long)W * a.H_total;
    const int gwarp = (int)blockIdx.x * K8_WARPS + warp;
    const int total_warps = (int)gridDim.x * K8_WARPS;
    float* const stage = DIRS == 2 ? a.stage_pts + (size_t)gwarp * (K8_DS * 384) : nullptr;
    uint32_t* const stage_vb = (DIRS == 2 && (a.pix || a.rgb)) ? a.stage_vb + (size_t)gwarp * (K8_DS * 4) : nullptr;

    int it = 0;                      // chunks decoded by this warp
    int rh = 0, rt = 0;              // pending ring: head (oldest), tail
    bool end_seen = false;           // this warp has drawn its one position past the end

    auto resolve_one = [&]() {
        const int2 e = ring[rh & (K8_RING - 1)];
        rh++;
        resolve8(a, e, tag, stage, stage_vb, n_chunks, lane);
    };
    // a chunk whose count is out joins the pending ring (resolved K8_D decoded chunks later)
    auto push = [&](int c, uint32_t n, int slot) {       // (callers make room first)
        if (lane == 0) ring[rt & (K8_RING - 1)] = make_int2(c, (int)(n | ((uint32_t)slot << 16) | ((uint32_t)(it & 0xff) << 24)));
        rt++;
        __syncwarp();
    };

    // ---- work: the next chunk with ROI pixels, or -1 = no more work, or -2 = the pending ring is full and cannot be
    //      emptied yet (the undecided position is kept in `carry`).  first = a position drawn earlier, or -1 to draw
    //      now; low = the lowest chunk this warp holds whose count is not out yet.  A pending entry above `low` must
    //      not be resolved here: its look-back would wait for this very warp. ----
    int carry = -1;
    auto next_chunk = [&](int first, int low) -> int {
        int c = first;
        for (;;) {
            if (DIRS == 2)
                while (rt - rh >= K8_RING - 1) {
                    if (ring[rh & (K8_RING - 1)].x > low) { carry = c; return -2; }
                    resolve_one();
                }
            if (c < 0) {
                if (end_seen) return -1;
                c = 0;
                if (lane == 0) c = (int)(atomicAdd(a.sched_ctr, 1u) - a.pos_base);
                c = __shfl_sync(0xffffffffu, c, 0);
            }
            if (c >= n_chunks) { end_seen = true; return -1; }
            const int p0 = c * 128, wc = min(128, plane - p0);
            const uint8_t* r = a.roi + (size_t)a.row0 * W + p0;
            // whoever draws chunk c pulls the ROI bytes of chunk c + (number of warps) into L2: about where the
            // draws will be one chunk period from now
            if (lane == 0 && c + total_warps < n_chunks) prefetch_l2(r + (size_t)total_warps * 128);
            uint32_t v = 0;
            if (4 * lane < wc) v = __ldg(reinterpret_cast<const uint32_t*>(r) + lane);
            if (__any_sync(0xffffffffu, v != 0)) return c;
            // no selected pixel: constant outputs (phase 0, fringe order -1, invalid, c_p_map 0), zero count
            if (4 * lane < wc) {
                const size_t g = (size_t)p0 + 4 * lane;
                *reinterpret_cast<float4*>(a.unw_v + g) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<uint2*>(a.code_v + g) = make_uint2(~0u, ~0u);
                *reinterpret_cast<uint32_t*>(a.valid + g) = 0u;
                if (DIRS == 2) {
                    *reinterpret_cast<float4*>(a.unw_h + g) = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<uint2*>(a.code_h + g) = make_uint2(~0u, ~0u);
                    reinterpret_cast<int4*>(a.cpmap + g)[0] = make_int4(0, 0, 0, 0);
                    reinterpret_cast<int4*>(a.cpmap + g)[1] = make_int4(0, 0, 0, 0);
                }
            }
            if (DIRS == 2) {
                if (lane == 0) st_state(a.tile_state + c, tag | ((c ? 1ull : 2ull) << 32));
                push(c, 0, 0);
            }
            c = -1;
        }
    };

    // ---- the warp's own loads: [NF frames] x [128 B] sub-tile + 4 ROI rows, all completing on the warp's mbarrier ----
    auto issue_load = [&](int c) {
        const int p0 = c * 128, wc = min(128, plane - p0);
        const long long gbase = (long long)a.row0 * W + p0 - ROI_HALO;
        long long seg0 = 0, seg1 = 0;
        uint32_t roi_tx = 0;
        if (lane < 4) {
            seg0 = max(gbase + (long long)(lane - 2) * W, 0LL);
            seg1 = min(gbase + (long long)(lane - 2) * W + wc + 2 * ROI_HALO, roi_total);
            if (seg1 > seg0) roi_tx = (uint32_t)(seg1 - seg0);
        }
        uint32_t roi_sum = roi_tx;
        roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 1);
        roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 2);
        roi_sum = __shfl_sync(0xffffffffu, roi_sum, 0);
        if (lane == 0) {
            fence_async_smem();      // the warp's reads of the sub-slot are done; the copy below rewrites it
            mbar_expect_tx(my_full, (uint32_t)NF * (a.use_tmap ? 128u : (uint32_t)wc) + roi_sum);
        }
        __syncwarp();
        const uint32_t dst = smem_u32(sw);
        if (a.use_tmap) {
            // one 2-D tensor copy (64-bit elements; bytes past the end of a frame are zero-filled): a dense [NF][128] block
            if (lane == 0) tensor_g2s_2d(dst, &stack_map, p0 >> 3, 0, my_full);
        } else {
            const uint8_t* src = a.stack + p0;
            for (int f = lane; f < NF; f += 32) bulk_g2s(dst + f * 128, src + (size_t)f * plane, (uint32_t)wc, my_full);
        }
        if (roi_tx)
            bulk_g2s(smem_u32(sroi) + lane * K8_ROI_ROW + (uint32_t)(seg0 - (gbase + (long long)(lane - 2) * W)),
                     a.roi + seg0, roi_tx, my_full);
    };

    constexpr int NONE_LOW = 0x7fffffff;
    int cur = next_chunk(-1, NONE_LOW);
    if (cur >= 0) issue_load(cur);
    int nxt = cur >= 0 ? next_chunk(-1, cur) : -1;

    while (cur >= 0) {
        // the draw for the chunk after next goes out now and is looked at after the FP64 phase
        int fut = -1;
        const bool fut_drawn = !end_seen && nxt != -2;      // (nxt == -2: an undecided position is carried, see below)
        if (fut_drawn && lane == 0) fut = (int)(atomicAdd(a.sched_ctr, 1u) - a.pos_base);

        if (!mbar_try(my_full, it & 1))
            while (!mbar_try(my_full, it & 1)) __nanosleep(64);
        const int p0 = cur * 128, wc = min(128, plane - p0);
        const int lpw = 4 * lane;
        const bool active = lpw < wc;
        const int row = (p0 + lpw) / W;
        const int xt = (p0 + lpw) - row * W;
        const int y = a.row0 + row;

        // ---------------- integer phase: mask, fringe terms, Gray bits ----------------
        uint32_t mbits = 0;
        Terms Tv, Th;
        uint32_t gvA = 0, gvB = 0, ghA = 0, ghB = 0;
        if (active) {
            const bool window_ok = y >= 2 && y + 1 < a.H_total && xt >= 4 && xt + 7 < W;
            bool fast = false;
            if (window_ok) {
                uint32_t any_zero = 0;
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(sroi + r * K8_ROI_ROW + (lpw + ROI_HALO - 4) + 4 * c);
                        any_zero |= (v - 0x01010101u) & ~v & 0x80808080u;
                    }
                if (any_zero == 0) { mbits = 0xf; fast = true; }
            }
            if (!fast) {
                const uint32_t centre = *reinterpret_cast<const uint32_t*>(sroi + 2 * K8_ROI_ROW + lpw + ROI_HALO);
                if (centre != 0) mbits = mask_slow8(sroi, lpw, xt, y, W, a.H_total);
            }
            if (mbits) {
                fringe_terms<N>(sw, 0, 32, lane, Tv);
                gray_bits(sw, N, N + a.M_v, a.M_v, 32, lane, gvA, gvB);
                if (DIRS == 2) {
                    const int fh = N + 2 * a.M_v;
                    fringe_terms<N>(sw, fh, 32, lane, Th);
                    gray_bits(sw, fh + N, fh + N + a.M_h, a.M_h, 32, lane, ghA, ghB);
                }
            }
        }
        // the sub-slot is consumed: the next chunk's frames come in while the FP64 phase runs
        __syncwarp();
        if (nxt >= 0) issue_load(nxt);

        // ---------------- FP64 phase (registers only) ----------------
        uint32_t vbits = 0;
        int4 cp01 = make_int4(0, 0, 0, 0), cp23 = cp01;   // the 4 pixels' correspondences, for the triangulation below
        const size_t g = (size_t)p0 + lpw;
        if (__any_sync(0xffffffffu, mbits != 0)) {
            if (active) {
                // Two passes of 2 pixels: inside a pass everything is straight-line (2 pixels x 2
                // directions interleave in the FP64 pipe); the pass loop is rolled to keep the
                // loop inside the instruction cache.
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    float r_unwv[2], r_unwh[2];
                    int r_cv[2], r_ch[2];
                    int2 r_cp[2];
                    uint32_t vb = 0;
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int j = 2 * h + u;
                        const int x = xt + j;
                        const bool m = (mbits >> j) & 1u;
                        const int cv = code_of(gvA, gvB, j, a.M_v);
                        const float wv = add_pi(phase_of<N>(Tv, j, tab));                    // 4/phase_unwrap.cpp:290
                        float unwv = (x == 0 || x == W - 1) ? 0.0f : unwrap_abs(wv, cv, fastdiv);  // :285, :291
                        unwv = m ? unwv : 0.0f;
                        bool v = m;
                        r_cv[u] = m ? cv : -1;
                        if (DIRS == 2) {
                            const int ch = code_of(ghA, ghB, j, a.M_h);
                            const float wh = add_pi(phase_of<N>(Th, j, tab));
                            float unwh = (y == 0 || y == a.H_total - 1) ? 0.0f : unwrap_abs(wh, ch, fastdiv);  // :304, :309
                            unwh = m ? unwh : 0.0f;
                            int px, py;                                                       // 5/compute_correspondance.cpp:648-675
                            const bool okx = correspond32(unwv, a.fw_v_d, &px);
                            const bool oky = correspond32(unwh, a.fw_h_d, &py);
                            // FE_INVALID on x rejects before y is computed (:650-655); on y after x is stored
                            r_cp[u].x = (m && okx) ? px : 0;
                            r_cp[u].y = (m && okx && oky) ? py : 0;
                            // 0 <= p <= P-1 as one unsigned compare (saturated values fall outside as well)
                            v = m && okx && oky && (unsigned)px <= (unsigned)(a.PW - 1) && (unsigned)py <= (unsigned)(a.PH - 1);
                            r_unwh[u] = unwh;
                            r_ch[u] = m ? ch : -1;
                        }
                        r_unwv[u] = unwv;
                        vb |= (v ? 1u : 0u) << u;
                    }
                    // plane outputs of the 2 pixels: one (vector) store per plane
                    const size_t gh = g + 2 * h;
                    *reinterpret_cast<float2*>(a.unw_v + gh) = make_float2(r_unwv[0], r_unwv[1]);
                    *reinterpret_cast<uint32_t*>(a.code_v + gh) = (uint32_t)(r_cv[0] & 0xffff) | ((uint32_t)r_cv[1] << 16);
                    *reinterpret_cast<uint16_t*>(a.valid + gh) = (uint16_t)((vb & 1u) | ((vb & 2u) << 7));
                    if (DIRS == 2) {
                        *reinterpret_cast<float2*>(a.unw_h + gh) = make_float2(r_unwh[0], r_unwh[1]);
                        *reinterpret_cast<uint32_t*>(a.code_h + gh) = (uint32_t)(r_ch[0] & 0xffff) | ((uint32_t)r_ch[1] << 16);
                        const int4 c4 = make_int4(r_cp[0].x, r_cp[0].y, r_cp[1].x, r_cp[1].y);
                        *reinterpret_cast<int4*>(a.cpmap + gh) = c4;
                        if (h == 0) cp01 = c4; else cp23 = c4;
                        // the undistortion tables are gathered once per surviving pixel a few thousand
                        // cycles from now: pull the lines into L2 meanwhile (DRAM latency -> L2 latency)
                        if (vb) {
                            if (a.cam_lut) prefetch_l2(a.cam_lut + gh);
                            if (a.proj_lut) {
                                if (vb & 1u) prefetch_l2(a.proj_lut + (size_t)c4.y * a.PW + c4.x);
                                if (vb & 2u) prefetch_l2(a.proj_lut + (size_t)c4.w * a.PW + c4.z);
                            }
                        }
                    }
                    vbits |= vb << (2 * h);
                }
            }
        } else if (active) {
            // the mask recurrence left no pixel of the chunk: constant outputs
            *reinterpret_cast<float4*>(a.unw_v + g) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<uint2*>(a.code_v + g) = make_uint2(~0u, ~0u);
            *reinterpret_cast<uint32_t*>(a.valid + g) = 0u;
            if (DIRS == 2) {
                *reinterpret_cast<float4*>(a.unw_h + g) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<uint2*>(a.code_h + g) = make_uint2(~0u, ~0u);
                reinterpret_cast<int4*>(a.cpmap + g)[0] = make_int4(0, 0, 0, 0);
                reinterpret_cast<int4*>(a.cpmap + g)[1] = make_int4(0, 0, 0, 0);
            }
        }
        if (DIRS == 2) {
            // ---- the chunk's survivors: the count goes out at once (every higher chunk's look-back needs it), the
            //      points are solved straight into a staging slot, their final place is found K8_D chunks from now ----
            const int slot = it % K8_DS;
            const uint32_t cnt = __popc(vbits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) st_state(a.tile_state + cur, tag | ((cur ? 1ull : 2ull) << 32) | wtotal);
            if (stage_vb) {
                // the threads' valid bits as 4 ballots, for the pixel indices / colours that go with the points
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (vbits >> j) & 1u);
                    if (lane == 0) stage_vb[slot * 4 + j] = m;
                }
            }
            float* cx = stage + slot * 384;
            uint32_t rank = incl - cnt;
            // triangulation of the surviving pixels (7/triangulation.cpp:1230-1247); the camera table entry of the
            // next pixel is in flight while this one is solved
            double2 lut_next = make_double2(0.0, 0.0);
            if (a.cam_lut && vbits) lut_next = a.cam_lut[g];
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                const double2 lut_cur = lut_next;
                if (a.cam_lut && vbits && j < 3) lut_next = a.cam_lut[g + j + 1];
                if (!((vbits >> j) & 1u)) continue;
                const int x = xt + j;
                const int cpx = j == 0 ? cp01.x : j == 1 ? cp01.z : j == 2 ? cp23.x : cp23.z;
                const int cpy = j == 0 ? cp01.y : j == 1 ? cp01.w : j == 2 ? cp23.y : cp23.w;
                double uc, vc, up, vp, Xd[3];
                if (a.cam_lut) {
                    uc = lut_cur.x; vc = lut_cur.y;
                } else {
                    undistorted_pixel_nodist(cal.Kc, cal.ifx_c, cal.ify_c, cal.cam_std != 0, (double)x, (double)y, &uc, &vc);
                }
                if (a.proj_lut) {
                    const double2 t = a.proj_lut[(size_t)cpy * a.PW + cpx];
                    up = t.x; vp = t.y;
                } else {
                    undistorted_pixel_nodist(cal.Kp, cal.ifx_p, cal.ify_p, cal.proj_std != 0, (double)cpx, (double)cpy, &up, &vp);
                }
                if (EXACT) triangulate_point(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                else triangulate_point_fast(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                cx[3 * rank + 0] = __double2float_rn(Xd[0]);                       // 8/save_point_cloud.cpp:94-96
                cx[3 * rank + 1] = __double2float_rn(Xd[1]);
                cx[3 * rank + 2] = __double2float_rn(Xd[2]);
                rank++;
            }
            __syncwarp();
            while (rt - rh >= K8_RING) resolve_one();       // (every pending entry is below cur, whose count is out)
            push(cur, wtotal, slot);
        }
        it++;

        if (nxt == -2) {
            // the pending ring was full of chunks above the one just decoded: it can be emptied now; the load of
            // the next chunk goes out late (a bubble, only after very long runs of chunks without ROI pixels)
            nxt = next_chunk(carry, NONE_LOW);
            if (nxt >= 0) issue_load(nxt);
        }
        // ---- the chunk after next: the draw made at the top (zero-count chunks on the way are dealt with there) ----
        int nxt2 = -1;
        if (nxt >= 0) nxt2 = next_chunk(fut_drawn ? __shfl_sync(0xffffffffu, fut, 0) : -1, nxt);

        // ---- pending chunks that are K8_D chunks old: find their place, move their points ----
        if (DIRS == 2)
            while (rh < rt && ((it - ((unsigned)ring[rh & (K8_RING - 1)].y >> 24)) & 0xff) >= K8_D) resolve_one();

        cur = nxt;
        nxt = nxt2;
    }
    if (DIRS == 2)
        while (rh < rt) resolve_one();
}

// 2-D view of the capture stack for the tile loads: inner dimension = one frame as 64-bit words,
// outer = the frames; box = [16 words = 128 B] x [NF frames], landing in shared memory as a dense [NF][128] block.
static bool stack_tensor_map8(CUtensorMap* map, const uint8_t* stack, size_t plane, int NF)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked_up = false;
    if (!looked_up) {
        looked_up = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = (encode_fn)fn;
    }
    if (!encode || NF > 256 || (plane & 15) || ((uintptr_t)stack & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)(plane / 8), (cuuint64_t)NF};
    const cuuint64_t gstride[1] = {(cuuint64_t)plane};
    const cuuint32_t box[2] = {16u, (cuuint32_t)NF};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<uint8_t*>(stack), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// words of staging a launch on sm_count SMs needs (points: floats; valid-bit ballots: uint32)
size_t fused8_stage_floats(int sm_count) { return (size_t)sm_count * S3D_K8_MINB * K8_WARPS * K8_DS * 384; }
size_t fused8_stage_vb_words(int sm_count) { return (size_t)sm_count * S3D_K8_MINB * K8_WARPS * K8_DS * 4; }

template <int N, int DIRS, int MINB, bool EXACT>
static cudaError_t launch8_t(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const Plan8& p, Fused8Cache* cache,
                             uint32_t* advance, cudaStream_t st)
{
    auto kern = k_fused8<N, DIRS, MINB, EXACT>;
    static size_t smem_cached = 0;      // per instantiation: attributes and occupancy are set once per process and size
    static int per_sm_cached = 0;
    if (smem_cached != p.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K8_WARPS * 32, p.smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (getenv("SCAN3D_DEBUG")) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            fprintf(stderr, "k_fused8<%d,%d,%d,%d>: occupancy %d CTAs/SM, %d regs, %zu B dyn smem, local %zu B\n", N, DIRS, MINB,
                    (int)EXACT, per_sm, fa.numRegs, p.smem, fa.localSizeBytes);
        }
        per_sm_cached = per_sm > MINB ? MINB : per_sm;
        smem_cached = p.smem;
    }
    const int want = (a.n_tiles + K8_WARPS - 1) / K8_WARPS;
    const int grid = want < sm_count * per_sm_cached ? want : sm_count * per_sm_cached;   // all CTAs resident
    FusedArgs a2 = a;
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    // the tensor map depends on the stack's address only: keep the last few (a ring of resident stacks is the common case)
    const CUtensorMap* map = nullptr;
    static const bool no_tmap = getenv("SCAN3D_NO_TMAP") != nullptr;
    if (!no_tmap) {
        for (int i = 0; i < Fused8Cache::SLOTS; i++)
            if (cache->stack[i] == a.stack) map = &cache->map[i];
        if (!map) {
            const int i = cache->next++ % Fused8Cache::SLOTS;
            cache->stack[i] = nullptr;
            if (stack_tensor_map8(&cache->map[i], a.stack, (size_t)a.W * a.H, NF)) {
                cache->stack[i] = a.stack;
                map = &cache->map[i];
            }
        }
    }
    alignas(64) CUtensorMap dummy;
    if (!map) { memset(&dummy, 0, sizeof(dummy)); map = &dummy; }
    a2.use_tmap = map != &dummy;
    kern<<<grid, K8_WARPS * 32, p.smem, st>>>(a2, cal, *map);
    *advance = (uint32_t)a.n_tiles + (uint32_t)grid * K8_WARPS;     // every warp draws one position past the end
    return cudaGetLastError();
}

template <int N, int DIRS>
static cudaError_t launch8_nd(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const Plan8& p, bool exact,
                              Fused8Cache* cache, uint32_t* advance, cudaStream_t st)
{
#define S3D_CASE8(MB)                                                                                      \
    if (p.minb == MB) {                                                                                    \
        if (exact || DIRS == 1) return launch8_t<N, DIRS, MB, true>(a, cal, sm_count, p, cache, advance, st); \
        return launch8_t<N, DIRS, MB, (DIRS == 1)>(a, cal, sm_count, p, cache, advance, st);                  \
    }
    S3D_CASE8(S3D_K8_MINB) S3D_CASE8(2)
#undef S3D_CASE8
    return cudaErrorInvalidValue;
}

cudaError_t launch_fused8(const scan3d_config& c, const FusedArgs& a_in, const DeviceCalib& cal, int sm_count,
                          Fused8Cache* cache, uint32_t* advance, cudaStream_t st)
{
    Plan8 p;
    if (!plan8(c, &p)) return cudaErrorInvalidValue;
    if (((uintptr_t)a_in.stack & 15) || ((uintptr_t)a_in.roi & 15)) return cudaErrorMisalignedAddress;
    if (c.dirs == 2 && !a_in.stage_pts) return cudaErrorInvalidValue;
    FusedArgs a = a_in;
    a.tiles_per_row = 0;
    a.n_tiles = fused8_num_chunks(c);
    a.dynamic = 1;
    const bool exact = !(c.flags & SCAN3D_FLAG_FAST_TRIANGULATION);
#define S3D_F8(NN) (c.dirs == 2 ? launch8_nd<NN, 2>(a, cal, sm_count, p, exact, cache, advance, st) : launch8_nd<NN, 1>(a, cal, sm_count, p, exact, cache, advance, st))
    switch (c.N) {
        case 3: return S3D_F8(3);
        case 4: return S3D_F8(4);
        case 5: return S3D_F8(5);
        case 8: return S3D_F8(8);
    }
#undef S3D_F8
    return cudaErrorInvalidValue;
}

}  // namespace s3d
