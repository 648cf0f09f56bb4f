// F + S: grid FAST detection and sub-pixel corner refinement.
//
// Reference behaviour replaced: FeatureExtractor::detectGridFAST
// (/root/reference/src/feature_extractor.cpp:443-570).  Defined semantics = the sequential
// ascending cell order (the reference's parallel_for_ body races on `mask`).
//
//   fast_cells_kernel   (F1, parallel over cells x frames)  FAST-9/16 score + 3x3 NMS inside each
//                       cell ROI exactly as cv::FastFeatureDetector(th, true, TYPE_9_16) on
//                       im(hroi) does (3-px dead border per cell), candidates in scan order,
//                       pre-filtered by the CV_32F-mask byte aliasing (cell-local x % 4 in {2,3}).
//   fast_sweep_kernel   (F2, one warp per frame)  the order-dependent part: occupancy, mask discs
//                       (cv::circle midpoint rasterisation) kept as a 1-bit/pixel bitmap in shared
//                       memory, per-cell mask filter -> libstdc++ std::sort top-1 -> response >= 20
//                       -> paint disc; adaptive nfast_th_ update (:546-552).
//   subpix_kernel       (S, one warp per point)  cv::cornerSubPix((3,3),(-1,-1),{30,0.01}).
//
// All integer work is bit-exact; S is float32/float64 with OpenCV's operation grouping
// (compiled with -fmad=false; the one FMA OpenCV uses is spelled __fmaf_rn).
#include "ov2_common.cuh"
#include "stdsort_emul.h"

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include <math.h>
#include <stdlib.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_CELL = 128;
constexpr int MAX_RADIUS = MAX_CELL / 4;

// FAST ring, clockwise from 12 o'clock
__constant__ int c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_ring_dy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};

struct FastArgs {
    const uint8_t* img; int w, h, pitch; long long fstride;   // level 0
    int first, cs, nwc, nhc, cap;
    const int32_t* th;          // [count] threshold per frame
    uint32_t* cand;             // [count][ncells][cap]  (resp << 16) | (ky << 8) | kx
    int32_t* cand_n;            // [count][ncells]
    int32_t* overflow;          // set to 1 if a cell had more than cap candidates
    int score_mode;             // 0: bisection on the 9-run bit test, 1: sliding-window min (sparse table)
};

// ---- TMA staging (sm_90+/sm_100a): one elected thread issues a single cp.async.bulk.tensor that lands
// the whole cell ROI (box BOXW x cs bytes of frame z) in shared memory and completes on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n"   /* suspend-time hint: sleep, do not spin (the spin was 7.6 % of the kernel's instructions) */
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

// ---------------------------------------------------------------------------------- F1
// roi pitch `rp`: cs on the plain-load path, the TMA box width (64 / 128) when the tile is staged by TMA
__global__ void fast_cells_kernel(FastArgs A, const __grid_constant__ CUtensorMap tmap, int use_tma, int rp) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int cs = A.cs;
    const uint8_t* roi = smem;                              // rp * cs (TMA destination: 128-byte aligned)
    const int roi_bytes = (rp * cs + 127) & ~127;
    const int cs2p = (cs * cs + 3) & ~3;                    // padded to whole words (the emission scans keep[] as words)
    uint8_t* sc = smem + roi_bytes;                         // cs*cs: score s (0 = not a corner at th)
    uint8_t* keep = sc + cs2p;                              // cs*cs: NMS survivor under the mask rule
    uint16_t* clist = (uint16_t*)(keep + cs2p);             // corners, (y << 8) | x cell-local (compacted)
    uint16_t* qlist = clist + (cs - 6) * (cs - 6);          // pixels that pass the compass pre-test (compacted)
    __shared__ int s_nq;
    __shared__ int s_ncorner;
    __shared__ __align__(8) uint64_t s_bar;
    const int cell = blockIdx.x, fr = blockIdx.y;
    const int r = cell / A.nwc, c = cell - r * A.nwc;
    const int x0 = c * cs, y0 = r * cs;
    int32_t* out_n = A.cand_n + (size_t)fr * (A.nwc * A.nhc) + cell;
    // cells the reference never searches (feature_extractor.cpp:510)
    if (!(x0 + cs < A.w - 1 && y0 + cs < A.h - 1)) {
        if (threadIdx.x == 0) *out_n = 0;
        return;
    }
    const int th = A.th[fr];
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, (uint32_t)(rp * cs));
            // measured on B200: a uint8 tiled box must START on a 16-byte boundary in the innermost
            // dimension (any other x faults with "illegal instruction"), so the box starts at x0 & ~15
            // and the cell's ROI begins (x0 & 15) bytes into each staged row.  Box columns beyond the
            // image are zero-filled and never read.
            tma_load_3d(smem, &tmap, &s_bar, x0 & ~15, y0, A.first + fr);
        }
        roi = smem + (x0 & 15);
        for (int i = threadIdx.x; i < cs2p; i += blockDim.x) { sc[i] = 0; keep[i] = 0; }
        __syncthreads();             // barrier init visible to every waiter
        mbar_wait(&s_bar, 0);
    } else {
        const uint8_t* img = A.img + A.fstride * (A.first + fr) + (size_t)y0 * A.pitch + x0;
        for (int i = threadIdx.x; i < cs2p; i += blockDim.x) {
            int yy = i / cs, xx = i - yy * cs;
            if (i < cs * cs) smem[yy * rp + xx] = __ldg(img + (size_t)yy * A.pitch + xx);
            sc[i] = 0;
            keep[i] = 0;
        }
    }
    if (threadIdx.x == 0) { s_ncorner = 0; s_nq = 0; }
    __syncthreads();
    const int in_w = cs - 6;
    // pass 1a (every interior pixel): compass pre-test, survivors compacted into qlist.  Every 9-arc of
    // the 16-ring contains two ADJACENT compass points {0,4,8,12}, so a corner needs two adjacent compass
    // pixels brighter than v + th (or darker than v - th).  Doing the full ring test here instead would
    // run it for nearly every warp (one passing lane is enough): ncu showed the 180-instruction ring block
    // executed by 92 % of the warp iterations.  (y, x) advance incrementally: no division per pixel.
    {
        const int npx = in_w * in_w, lane = threadIdx.x & 31;
        int i = threadIdx.x;
        int yy = i / in_w, xx = i - yy * in_w;
        const int sy = (int)blockDim.x / in_w, sx = (int)blockDim.x - sy * in_w;
        for (int base = 0; base < npx; base += blockDim.x) {      // uniform trip count: ballots below
            bool pass = false;
            if (i < npx) {
                const uint8_t* p = roi + (yy + 3) * rp + (xx + 3);
                const int v = *p;
                const int c0 = v - (int)p[-3 * rp], c4 = v - (int)p[3], c8 = v - (int)p[3 * rp], c12 = v - (int)p[-3];
                const bool b0 = c0 > th, b4 = c4 > th, b8 = c8 > th, b12 = c12 > th;
                const bool d0 = c0 < -th, d4 = c4 < -th, d8 = c8 < -th, d12 = c12 < -th;
                pass = ((b0 || b8) && (b4 || b12)) || ((d0 || d8) && (d4 || d12));   // two adjacent compass points
            }
            const unsigned m = __ballot_sync(FULL, pass);
            if (m) {
                const int leader = __ffs(m) - 1;
                int slot = 0;
                if (lane == leader) slot = atomicAdd(&s_nq, __popc(m));
                slot = __shfl_sync(FULL, slot, leader);
                if (pass) qlist[slot + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(((yy + 3) << 8) | (xx + 3));
            }
            i += blockDim.x;
            xx += sx; yy += sy;
            if (xx >= in_w) { xx -= in_w; yy++; }
        }
    }
    __syncthreads();
    // ring offsets for this pitch, once per thread (they were 16 IMADs per tested pixel)
    int roff[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) roff[k] = c_ring_dy[k] * rp + c_ring_dx[k];
    // pass 1b (dense over the survivors): the 16-pixel ring test
    const int nq = s_nq;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int yx = qlist[q];
        const uint8_t* p = roi + (yx >> 8) * rp + (yx & 255);
        const int v = *p;
        const int vlo = v - th, vhi = v + th;              // bright: ring < v - th, dark: ring > v + th
        unsigned bright = 0, dark = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int rk = (int)p[roff[k]];
            bright |= (unsigned)(rk < vlo) << k;
            dark |= (unsigned)(rk > vhi) << k;
        }
        unsigned b = bright | (bright << 16), dk = dark | (dark << 16);
        b &= b >> 1; b &= b >> 2; b &= b >> 4; b &= b >> 1;
        dk &= dk >> 1; dk &= dk >> 2; dk &= dk >> 4; dk &= dk >> 1;
        if ((b | dk) & 0xFFFFu) {
            int pos = atomicAdd(&s_ncorner, 1);
            clist[pos] = (uint16_t)yx;
        }
    }
    __syncthreads();
    // pass 2 (dense over the corners): exact score s = max over the 16 arcs of min(d) / min(-d)
    // = the largest t for which a 9-run of (d >= t) or (d <= -t) exists: bisection on t.
    const int ncorner = s_ncorner;
    for (int i = threadIdx.x; i < ncorner; i += blockDim.x) {
        const int yx = clist[i];
        const int pi = (yx >> 8) * cs + (yx & 255);
        const uint8_t* p = roi + (yx >> 8) * rp + (yx & 255);
        const int v = *p;
        int d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = v - (int)p[roff[k]];
        int lo;
        if (A.score_mode == 1) {
            // s = max_k min(d[k..k+8]) (bright) / max_k min(-d[k..k+8]) (dark) with a sparse table:
            // m2 = windows of 2, m4 of 4, m8 of 8, then one more element -> 9.  inline PTX min/max keeps
            // this as written (a naive 16x8 min/max nest was folded into VIMNMX3 chains
            // that returned max(d)).
            int best = 0;
#pragma unroll
            for (int pol = 0; pol < 2; ++pol) {
                int e[16], m2[16], m4[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = pol ? -d[k] : d[k];
#pragma unroll
                for (int k = 0; k < 16; ++k) asm("min.s32 %0, %1, %2;" : "=r"(m2[k]) : "r"(e[k]), "r"(e[(k + 1) & 15]));
#pragma unroll
                for (int k = 0; k < 16; ++k) asm("min.s32 %0, %1, %2;" : "=r"(m4[k]) : "r"(m2[k]), "r"(m2[(k + 2) & 15]));
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    int m8, m9;
                    asm("min.s32 %0, %1, %2;" : "=r"(m8) : "r"(m4[k]), "r"(m4[(k + 4) & 15]));
                    asm("min.s32 %0, %1, %2;" : "=r"(m9) : "r"(m8), "r"(e[(k + 8) & 15]));
                    asm("max.s32 %0, %1, %2;" : "=r"(best) : "r"(best), "r"(m9));
                }
            }
            lo = best;
        } else {
        lo = th + 1;
        int hi = 256;   // P(lo) holds (corner at th), P(256) cannot
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            unsigned bm = 0, dm = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                bm |= (unsigned)(d[k] >= mid) << k;
                dm |= (unsigned)(d[k] <= -mid) << k;
            }
            bm |= bm << 16; dm |= dm << 16;
            bm &= bm >> 1; bm &= bm >> 2; bm &= bm >> 4; bm &= bm >> 1;
            dm &= dm >> 1; dm &= dm >> 2; dm &= dm >> 4; dm &= dm >> 1;
            if ((bm | dm) & 0xFFFFu) lo = mid; else hi = mid;
        }
        }
        sc[pi] = (uint8_t)lo;   // th < s <= 255
    }
    __syncthreads();
    // 3x3 NMS on response = s - 1 (non-corners count 0): strictly greater than all 8 neighbours
    for (int i = threadIdx.x; i < ncorner; i += blockDim.x) {
        const int yx = clist[i];
        const int xx = yx & 255, pi = (yx >> 8) * cs + xx;
        const int resp = (int)sc[pi] - 1;
        bool ok = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0) continue;
                int sn = sc[pi + dy * cs + dx];
                int rn = sn ? sn - 1 : 0;
                ok = ok && (resp > rn);
            }
        // CV_32F mask read as bytes: only cell-local x % 4 in {2,3} can see a non-zero byte of 1.0f
        if (ok && (xx & 2)) keep[pi] = 1;
    }
    __syncthreads();
    // ordered emission (row-major scan order) by warp 0: keep[] is scanned as 32-bit words (almost all
    // zero); the few non-zero words are expanded serially, so the output order is the scan order
    if (threadIdx.x < 32) {
        uint32_t* out = A.cand + ((size_t)fr * (A.nwc * A.nhc) + cell) * A.cap;
        const uint32_t* kw = reinterpret_cast<const uint32_t*>(keep);
        const int nw = cs2p >> 2, lane = threadIdx.x;
        const unsigned rcs = 0xFFFFFFFFu / (unsigned)cs + 1u;          // ceil(2^32 / cs): floor(pi * rcs / 2^32) == pi / cs for pi < 2^16
        int n = 0;
        for (int base = 0; base < nw; base += 32) {
            const int wi = base + lane;
            const uint32_t w = wi < nw ? kw[wi] : 0u;
            unsigned m = __ballot_sync(FULL, w != 0u);
            while (m) {                                    // uniform
                const int L = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t ww = __shfl_sync(FULL, w, L);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if ((ww >> (8 * b)) & 0xFFu) {
                        if (lane == 0 && n < A.cap) {
                            const int pi = (base + L) * 4 + b;
                            const int yy = (int)__umulhi((unsigned)pi, rcs), xx = pi - yy * cs;   // pi / cs, exact for pi < 2^16
                            out[n] = ((uint32_t)(sc[pi] - 1) << 16) | ((uint32_t)yy << 8) | (uint32_t)xx;
                        }
                        n++;
                    }
                }
            }
        }
        if (lane == 0) {
            if (n > A.cap) { atomicExch(A.overflow, 1); n = A.cap; }
            *out_n = n;
        }
    }
}

// ---------------------------------------------------------------------------------- F2
struct SweepArgs {
    int w, h, cs, nwc, nhc, cap, radius, max_per_frame;
    int hw[MAX_RADIUS + 1];          // cv::circle half-widths per |dy|
    const int32_t* kp_off;           // [count+1] or NULL
    const float2* kps;               // existing keypoints
    const uint32_t* cand; const int32_t* cand_n;
    int32_t* th;                     // in/out per frame
    int2* out_int;                   // [count][max_per_frame]
    int32_t* out_n;                  // [count]
};

__device__ __forceinline__ void paint_row(uint32_t* bm, int wpr, int W, int H, int y, int xa, int xb) {
    if (y < 0 || y >= H) return;
    xa = max(xa, 0);
    xb = min(xb, W - 1);
    if (xa > xb) return;
    uint32_t* row = bm + (size_t)y * wpr;
    int wa = xa >> 5, wb = xb >> 5;
    for (int wi = wa; wi <= wb; ++wi) {
        uint32_t m = 0xffffffffu;
        if (wi == wa) m &= 0xffffffffu << (xa & 31);
        if (wi == wb) m &= 0xffffffffu >> (31 - (xb & 31));
        atomicOr(row + wi, m);
    }
}

__global__ void __launch_bounds__(32) fast_sweep_kernel(SweepArgs A) {
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x, fr = blockIdx.x;
    const int wpr = (A.w + 31) >> 5;
    uint32_t* bm = sm;                                  // h * wpr words, bit = 1: mask is 0.0f there
    int* skey = (int*)(sm + (size_t)A.h * wpr);         // cap sort keys
    uint8_t* occ = (uint8_t*)(skey + A.cap);            // (nhc+1)*(nwc+1)
    for (int i = lane; i < A.h * wpr; i += 32) bm[i] = 0;
    const int nocc = (A.nhc + 1) * (A.nwc + 1);
    for (int i = lane; i < nocc; i += 32) occ[i] = 0;
    __syncwarp();
    // existing keypoints: occupancy + discs (feature_extractor.cpp:471-474)
    if (A.kp_off) {
        const int k0 = A.kp_off[fr], k1 = A.kp_off[fr + 1];
        const int nrows = 2 * A.radius + 1;
        for (int k = k0; k < k1; ++k) {
            float2 px = A.kps[k];
            if (lane == 0) {
                int rr = (int)(px.y / (float)A.cs), cc = (int)(px.x / (float)A.cs);
                if (rr >= 0 && rr <= A.nhc && cc >= 0 && cc <= A.nwc) occ[rr * (A.nwc + 1) + cc] = 1;
            }
            int cx = __float2int_rn(px.x), cy = __float2int_rn(px.y);
            for (int rI = lane; rI < nrows; rI += 32) {
                int dy = rI - A.radius;
                int hwv = A.hw[dy < 0 ? -dy : dy];
                paint_row(bm, wpr, A.w, A.h, cy + dy, cx - hwv, cx + hwv);
            }
        }
    }
    __syncwarp();
    const int ncells = A.nwc * A.nhc;
    const uint32_t* cand = A.cand + (size_t)fr * ncells * A.cap;
    const int32_t* cand_n = A.cand_n + (size_t)fr * ncells;
    int2* out = A.out_int + (size_t)fr * A.max_per_frame;
    int nbkps = 0, nbempty = 0;
    for (int cell = 0; cell < ncells; ++cell) {
        const int r = cell / A.nwc, c = cell - r * A.nwc;
        if (occ[r * (A.nwc + 1) + c]) continue;
        nbempty++;
        const int x0 = c * A.cs, y0 = r * A.cs;
        if (!(x0 + A.cs < A.w - 1 && y0 + A.cs < A.h - 1)) continue;
        const int n = cand_n[cell];
        if (n == 0) continue;
        // mask filter + ordered compaction of survivors into skey
        int ns = 0, vmax = -1;
        for (int base = 0; base < n; base += 32) {
            int j = base + lane;
            bool ok = false;
            uint32_t cd = 0;
            if (j < n) {
                cd = cand[(size_t)cell * A.cap + j];
                int kx = cd & 255, ky = (cd >> 8) & 255;
                int mx = x0 + (kx >> 2), my = y0 + ky;
                ok = ((bm[(size_t)my * wpr + (mx >> 5)] >> (mx & 31)) & 1u) == 0;
            }
            unsigned m = __ballot_sync(FULL, ok);
            if (ok) {
                int pos = ns + __popc(m & ((1u << lane) - 1));
                skey[pos] = (int)((cd >> 16) << 8) | j;   // (response << 8) | candidate index
                vmax = max(vmax, (int)(cd >> 16));
            }
            ns += __popc(m);
        }
        if (ns == 0) continue;
        vmax = __reduce_max_sync(FULL, vmax);
        __syncwarp();
        // number of survivors sharing the maximal response, and the first of them in scan order
        int nmax = 0, firstmax = 0x7fffffff;
        for (int base = 0; base < ns; base += 32) {
            int j = base + lane;
            bool is = j < ns && (skey[j] >> 8) == vmax;
            unsigned m = __ballot_sync(FULL, is);
            if (m && firstmax == 0x7fffffff) firstmax = base + __ffs(m) - 1;
            nmax += __popc(m);
        }
        int win_j;
        if (nmax == 1 || ns <= 16) {
            // unique maximum, or libstdc++ falls straight to (stable) insertion sort
            win_j = skey[firstmax] & 255;
        } else {
            if (lane == 0) ov2sort::sort_desc(skey, ns);
            __syncwarp();
            win_j = skey[0] & 255;
        }
        __syncwarp();
        if (vmax >= 20) {   // vkps.at(0).response >= 20
            uint32_t cd = cand[(size_t)cell * A.cap + win_j];
            int px = x0 + (int)(cd & 255), py = y0 + (int)((cd >> 8) & 255);
            if (lane == 0 && nbkps < A.max_per_frame) out[nbkps] = make_int2(px, py);
            nbkps++;
            const int nrows = 2 * A.radius + 1;
            for (int rI = lane; rI < nrows; rI += 32) {
                int dy = rI - A.radius;
                int hwv = A.hw[dy < 0 ? -dy : dy];
                paint_row(bm, wpr, A.w, A.h, py + dy, px - hwv, px + hwv);
            }
            __syncwarp();
        }
    }
    for (int i = nbkps + lane; i < A.max_per_frame; i += 32) out[i] = make_int2(-1, -1);
    if (lane == 0) {
        A.out_n[fr] = nbkps;
        int th = A.th[fr];
        if ((double)nbkps < 0.5 * (double)nbempty && nbempty > 10) th = (int)((double)th * 0.66);
        else if (nbkps == nbempty) th = (int)((double)th * 1.5);
        A.th[fr] = th;
    }
}

// Wavefront variant: the sequential cell order is the semantics, but a cell only reads mask bits of its own area, and a disc
// (radius cs / 4) reaches no further than the neighbouring cells - so cell (r, c) needs (r, c-1), (r-1, c-1), (r-1, c) and
// (r-1, c+1) finished, nothing else.  One warp per cell row, row r trailing row r-1 by two cells (as ss_sweep_kernel does):
// exactly the sequential result in nwc + 2 nhc cell times instead of nwc * nhc (one warp per frame was 10 % of the C2 step,
// pure latency).  Detections are kept per cell and put in cell order at the end.
constexpr int FSWEEP_MAX_WARPS = 16;

__global__ void __launch_bounds__(FSWEEP_MAX_WARPS * 32, 1) fast_sweep_wave_kernel(SweepArgs A) {
    extern __shared__ uint32_t sm[];
    __shared__ int s_cnt[2];
    const int tid = threadIdx.x, fr = blockIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int wpr = (A.w + 31) >> 5, ncells = A.nwc * A.nhc;
    uint32_t* bm = sm;                                                        // h * wpr words, bit = 1: mask is 0.0f there
    int2* det = reinterpret_cast<int2*>(sm + (((size_t)A.h * wpr + 1) & ~(size_t)1));   // per cell, x < 0: none
    int* skey_all = reinterpret_cast<int*>(det + ncells);                    // [nwarps][cap] sort keys
    volatile int* progress = reinterpret_cast<volatile int*>(skey_all + (size_t)nwarps * A.cap);   // [nhc]
    uint8_t* occ = reinterpret_cast<uint8_t*>(const_cast<int*>(progress) + A.nhc);       // (nhc+1)*(nwc+1)
    int* skey = skey_all + (size_t)warp * A.cap;
    for (int i = tid; i < A.h * wpr; i += nthreads) bm[i] = 0;
    for (int i = tid; i < ncells; i += nthreads) det[i] = make_int2(-1, -1);
    for (int i = tid; i < A.nhc; i += nthreads) progress[i] = 0;
    const int nocc = (A.nhc + 1) * (A.nwc + 1);
    for (int i = tid; i < nocc; i += nthreads) occ[i] = 0;
    __syncthreads();
    const int nrows = 2 * A.radius + 1;
    // existing keypoints: occupancy + discs (feature_extractor.cpp:471-474); painting commutes (atomicOr)
    if (A.kp_off) {
        const int k0 = A.kp_off[fr], k1 = A.kp_off[fr + 1];
        for (int k = k0 + tid; k < k1; k += nthreads) {
            const float2 px = A.kps[k];
            const int rr = (int)(px.y / (float)A.cs), cc = (int)(px.x / (float)A.cs);
            if (rr >= 0 && rr <= A.nhc && cc >= 0 && cc <= A.nwc) occ[rr * (A.nwc + 1) + cc] = 1;
            const int cx = __float2int_rn(px.x), cy = __float2int_rn(px.y);
            for (int rI = 0; rI < nrows; ++rI) {
                const int dy = rI - A.radius;
                const int hwv = A.hw[dy < 0 ? -dy : dy];
                paint_row(bm, wpr, A.w, A.h, cy + dy, cx - hwv, cx + hwv);
            }
        }
    }
    __syncthreads();
    const uint32_t* cand = A.cand + (size_t)fr * ncells * A.cap;
    const int32_t* cand_n = A.cand_n + (size_t)fr * ncells;
    for (int r = warp; r < A.nhc; r += nwarps) {
        for (int c = 0; c < A.nwc; ++c) {
            const int cell = r * A.nwc + c;
            const int x0 = c * A.cs, y0 = r * A.cs;
            const bool search = !occ[r * (A.nwc + 1) + c] && (x0 + A.cs < A.w - 1 && y0 + A.cs < A.h - 1);
            // the cell's candidates do not depend on any other cell: fetch them before waiting for the neighbours
            const int n = search ? cand_n[cell] : 0;
            uint32_t cdv[4] = {0u, 0u, 0u, 0u};                                // cap <= 128: four candidates per lane
            if (n > 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = 32 * q + lane;
                    if (j < n) cdv[q] = __ldg(cand + (size_t)cell * A.cap + j);
                }
            }
            if (r > 0) {
                const int need = min(c + 2, A.nwc);                            // row r-1 has finished cell c+1 (or its last cell)
                while (progress[r - 1] < need) __nanosleep(40);
                __threadfence_block();
            }
            if (n > 0) {
                // mask filter + ordered compaction of survivors into skey
                int ns = 0, vmax = -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = 32 * q + lane;
                    if (32 * q >= n) break;                                    // uniform
                    bool ok = false;
                    const uint32_t cd = cdv[q];
                    if (j < n) {
                        const int kx = cd & 255, ky = (cd >> 8) & 255;
                        const int mx = x0 + (kx >> 2), my = y0 + ky;
                        ok = ((bm[(size_t)my * wpr + (mx >> 5)] >> (mx & 31)) & 1u) == 0;
                    }
                    const unsigned m = __ballot_sync(FULL, ok);
                    if (ok) {
                        const int pos = ns + __popc(m & ((1u << lane) - 1));
                        skey[pos] = (int)((cd >> 16) << 8) | j;                // (response << 8) | candidate index
                        vmax = max(vmax, (int)(cd >> 16));
                    }
                    ns += __popc(m);
                }
                if (ns > 0) {
                    vmax = __reduce_max_sync(FULL, vmax);
                    __syncwarp();
                    int nmax = 0, firstmax = 0x7fffffff;
                    for (int base = 0; base < ns; base += 32) {
                        const int j = base + lane;
                        const bool is = j < ns && (skey[j] >> 8) == vmax;
                        const unsigned m = __ballot_sync(FULL, is);
                        if (m && firstmax == 0x7fffffff) firstmax = base + __ffs(m) - 1;
                        nmax += __popc(m);
                    }
                    int win_j;
                    if (nmax == 1 || ns <= 16) {
                        win_j = skey[firstmax] & 255;                          // unique maximum / libstdc++'s insertion-sort range
                    } else {
                        if (lane == 0) ov2sort::sort_desc(skey, ns);
                        __syncwarp();
                        win_j = skey[0] & 255;
                    }
                    __syncwarp();
                    if (vmax >= 20) {                                          // vkps.at(0).response >= 20
                        const int wq = win_j >> 5;                              // uniform
                        const uint32_t mine = wq == 0 ? cdv[0] : (wq == 1 ? cdv[1] : (wq == 2 ? cdv[2] : cdv[3]));
                        const uint32_t cd = __shfl_sync(FULL, mine, win_j & 31);
                        const int px = x0 + (int)(cd & 255), py = y0 + (int)((cd >> 8) & 255);
                        if (lane == 0) det[cell] = make_int2(px, py);
                        for (int rI = lane; rI < nrows; rI += 32) {
                            const int dy = rI - A.radius;
                            const int hwv = A.hw[dy < 0 ? -dy : dy];
                            paint_row(bm, wpr, A.w, A.h, py + dy, px - hwv, px + hwv);
                        }
                        __syncwarp();
                    }
                }
            }
            __threadfence_block();                                            // discs before the counter
            __syncwarp();
            if (lane == 0) progress[r] = c + 1;
        }
    }
    __syncthreads();
    // assembly in cell order + the adaptive threshold (feature_extractor.cpp:546-552)
    int2* out = A.out_int + (size_t)fr * A.max_per_frame;
    if (warp == 0) {
        int nbkps = 0, nbempty = 0;
        for (int base = 0; base < ncells; base += 32) {
            const int cell = base + lane;
            bool empty = false, has = false;
            int2 d = make_int2(-1, -1);
            if (cell < ncells) {
                const int r = cell / A.nwc, c = cell - r * A.nwc;
                empty = !occ[r * (A.nwc + 1) + c];
                d = det[cell];
                has = d.x >= 0;
            }
            const unsigned mh = __ballot_sync(FULL, has);
            if (has) {
                const int pos = nbkps + __popc(mh & ((1u << lane) - 1));
                if (pos < A.max_per_frame) out[pos] = d;
            }
            nbkps += __popc(mh);
            nbempty += __popc(__ballot_sync(FULL, empty));
        }
        if (lane == 0) {
            A.out_n[fr] = nbkps;
            int th = A.th[fr];
            if ((double)nbkps < 0.5 * (double)nbempty && nbempty > 10) th = (int)((double)th * 0.66);
            else if (nbkps == nbempty) th = (int)((double)th * 1.5);
            A.th[fr] = th;
            s_cnt[0] = nbkps;
        }
    }
    __syncthreads();
    for (int i = s_cnt[0] + tid; i < A.max_per_frame; i += nthreads) out[i] = make_int2(-1, -1);
}

// ---------------------------------------------------------------------------------- S
struct SubpixArgs {
    const uint8_t* img; int w, h, pitch; long long fstride;
    int first, max_per_frame, n;   // n = count * max_per_frame
    const int2* in_int;
    float2* out;
    int do_subpix;
    float mask[49];                // exp(-(i/3)^2) * exp(-(j/3)^2), computed on the host with libm expf
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// cv::getRectSubPix(8u -> 32f, 9x9) value at patch (i, j); grouping found by probing cv2 4.13
// (oracle/image_ref.py::_get_rect_subpix_9).
__device__ __forceinline__ float rect_subpix_px(const uint8_t* img, int pitch, int W, int H, int ix, int iy, float a,
                                                float b, int i, int j) {
    const bool inside = 0 <= ix && ix + 9 < W && 0 <= iy && iy + 9 < H;
    const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    if (inside) {
        const uint8_t* p = img + (size_t)(iy + i) * pitch + ix + j;
        float p00 = p[0], p01 = p[1], p10 = p[pitch], p11 = p[pitch + 1];
        return (p00 * a11 + p01 * a12) + (p10 * a21 + p11 * a22);
    }
    const int rx = ix < 0 ? min(-ix, 9) : 0;
    const int rw = ix < W - 9 ? 9 : max(W - ix - 1, 0);
    const int ry = iy < 0 ? -iy : 0;
    const int rh = iy < H - 9 ? 9 : max(H - iy - 1, 0);
    const int r0 = clampi(iy + i, 0, H - 1), r1 = clampi(iy + i + 1, 0, H - 1);
    const uint8_t* s0 = img + (size_t)r0 * pitch;
    const uint8_t* s1 = img + (size_t)r1 * pitch;
    if (j < rx || j >= rw) {
        int cc = j < rx ? 0 : (i < ry ? W - 2 : W - 1);
        cc = clampi(cc, 0, W - 1);
        return (float)s0[cc] * (1.f - b) + (float)s1[cc] * b;
    }
    const int cc = ix + j;
    if (i < ry || i >= rh) return __fmaf_rn((float)s0[cc + 1], a, (float)s0[cc] * (1.f - a));
    return ((float)s0[cc] * a11 + (float)s0[cc + 1] * a12) + ((float)s1[cc] * a21 + (float)s1[cc + 1] * a22);
}

__global__ void __launch_bounds__(128) subpix_kernel(SubpixArgs A) {
    __shared__ float spatch[4][81];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 4 + warp;
    if (idx >= A.n) return;
    const int2 ip = A.in_int[idx];
    if (ip.x < 0) {
        if (lane == 0) A.out[idx] = make_float2(-1.f, -1.f);
        return;
    }
    const float2 cT = make_float2((float)ip.x, (float)ip.y);
    if (!A.do_subpix) {
        if (lane == 0) A.out[idx] = cT;
        return;
    }
    const int fr = idx / A.max_per_frame;
    const uint8_t* img = A.img + A.fstride * (A.first + fr);
    float* sp = spatch[warp];
    float2 cI = cT;
    const double eps = 0.01 * 0.01;
    int iter = 0;
    double err = 0.0;
    do {
        const float x = cI.x - 4.f, y = cI.y - 4.f;
        const int ix = __float2int_rd(x), iy = __float2int_rd(y);
        const float a = x - (float)ix, b = y - (float)iy;
        __syncwarp();
        for (int p = lane; p < 81; p += 32) sp[p] = rect_subpix_px(img, A.pitch, A.w, A.h, ix, iy, a, b, p / 9, p % 9);
        __syncwarp();
        double sa = 0, sb = 0, scc = 0, sbb1 = 0, sbb2 = 0;
        for (int p = lane; p < 49; p += 32) {
            int i = p / 7, j = p - i * 7;
            const float* q = sp + (i + 1) * 9 + (j + 1);
            double m = (double)A.mask[p];
            double tgx = (double)(q[1] - q[-1]);
            double tgy = (double)(q[9] - q[-9]);
            double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
            double px = (double)(j - 3), py = (double)(i - 3);
            sa += gxx; sb += gxy; scc += gyy;
            sbb1 += gxx * px + gxy * py;
            sbb2 += gxy * px + gyy * py;
        }
        sa = warp_sum_d(sa); sb = warp_sum_d(sb); scc = warp_sum_d(scc);
        sbb1 = warp_sum_d(sbb1); sbb2 = warp_sum_d(sbb2);
        const double det = sa * scc - sb * sb;
        if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
        const double scale = 1.0 / det;
        float2 cI2;
        cI2.x = (float)((double)cI.x + scc * scale * sbb1 - sb * scale * sbb2);
        cI2.y = (float)((double)cI.y - sb * scale * sbb1 + sa * scale * sbb2);
        const float ex = cI2.x - cI.x, ey = cI2.y - cI.y;
        err = (double)(ex * ex + ey * ey);
        if (cI2.x < 0 || cI2.x >= (float)A.w || cI2.y < 0 || cI2.y >= (float)A.h) break;
        cI = cI2;
    } while (++iter < 30 && err > eps);
    if (fabsf(cI.x - cT.x) > 3.f || fabsf(cI.y - cT.y) > 3.f) cI = cT;
    if (lane == 0) A.out[idx] = cI;
}

// OpenCV's filled-circle rasterisation (midpoint algorithm): half-width per |dy|
void circle_halfwidths(int radius, int* hw) {
    for (int i = 0; i <= radius; ++i) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        if (dx > hw[dy]) hw[dy] = dx;
        if (dy > hw[dx]) hw[dx] = dy;
        dy++;
        err += plus;
        plus += 2;
        int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

}  // namespace

// Host side of the TMA staging: a 3-D tiled tensor map {W, H, frames} over the level-0 images, box
// {rp, cs, 1} (rp = cs + 15 rounded up to 16 bytes: the box starts on the 16-byte boundary at or below the cell).  Falls back to plain loads
// when the images do not meet the TMA alignment rules (user-aliased level 0 with an odd pitch).
typedef CUresult (*ov2_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ov2_encode_tiled_fn tma_encoder() {
    static ov2_encode_tiled_fn fn = []() -> ov2_encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (ov2_encode_tiled_fn)p;
    }();
    return fn;
}

static ov2_status launch_fast_cells(ov2_ctx* ctx, const ov2_pyr* pyr, const FastArgs& FA, int nframes_grid) {
    const int cs = FA.cs;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tma = 0, rp = cs;
    const char* env = getenv("OV2_NO_TMA");
    ov2_encode_tiled_fn enc = (env && atoi(env)) ? nullptr : tma_encoder();
    if (enc && ((uintptr_t)FA.img % 16) == 0 && FA.pitch % 16 == 0 && FA.fstride % 16 == 0) {
        const int box_w = (cs + 15 + 15) & ~15;
        cuuint64_t dims[3] = {(cuuint64_t)FA.w, (cuuint64_t)FA.h, (cuuint64_t)pyr->batch};
        cuuint64_t strides[2] = {(cuuint64_t)FA.pitch, (cuuint64_t)FA.fstride};
        cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)cs, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)FA.img, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) { use_tma = 1; rp = box_w; }
    }
    const int threads = cs > 24 ? 128 : 32;
    const size_t roi_bytes = ((size_t)rp * cs + 127) & ~(size_t)127;
    const size_t smem = roi_bytes + (size_t)2 * ((cs * cs + 3) & ~3) + (size_t)4 * (cs - 6) * (cs - 6);
    if (smem > 48 * 1024)
        OV2_CUDA(ctx, cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OV2_LAUNCH(ctx, "fast_cells_kernel",
               fast_cells_kernel<<<dim3(FA.nwc * FA.nhc, nframes_grid), threads, smem, ctx->stream>>>(FA, tmap, use_tma, rp));
    return OV2_OK;
}


extern "C" ov2_status ov2_grid_fast(ov2_ctx* ctx, const ov2_pyr* pyr, int first, int count, int cellsize,
                                    const int32_t* curkp_offsets, const float* curkps, int32_t* fast_th_inout,
                                    int max_per_frame, float* out_pts, int32_t* out_counts, int32_t* out_pts_int,
                                    int do_subpix) {
    if (!ctx || !pyr || !pyr->l0 || first < 0 || count <= 0 || first + count > pyr->batch || !fast_th_inout ||
        !out_pts || !out_counts)
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_grid_fast: bad arguments");
    if (cellsize < 8 || cellsize > MAX_CELL)
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_grid_fast: cellsize must be in [8, 128]");
    const int W = pyr->w[0], H = pyr->h[0];
    const int nwc = W / cellsize, nhc = H / cellsize, ncells = nwc * nhc;
    if (ncells == 0 || max_per_frame < ncells)
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_grid_fast: max_per_frame < number of cells");
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    const int in_w = cellsize - 6;
    int cap = ((in_w + 1) / 2) * ((in_w + 1) / 2);
    if (cap > 128) cap = 128;
    if (cap < 4) cap = 4;

    const void* d = nullptr;
    void* o = nullptr;
    int nkp_total = 0;
    const int32_t* d_off = nullptr;
    const float2* d_kps = nullptr;
    if (curkp_offsets) {
        if ((st = ov2_stage_in(ctx, curkp_offsets, sizeof(int32_t) * (size_t)(count + 1), &d)) != OV2_OK) return st;
        d_off = (const int32_t*)d;
        // number of keypoints: need offsets[count] on the host
        if (ov2_is_device_ptr(curkp_offsets)) {
            OV2_CUDA(ctx, cudaMemcpyAsync(&nkp_total, curkp_offsets + count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            OV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            nkp_total = curkp_offsets[count];
        }
        if (nkp_total > 0) {
            if (!curkps) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_grid_fast: curkps is NULL");
            if ((st = ov2_stage_in(ctx, curkps, sizeof(float) * 2 * (size_t)nkp_total, &d)) != OV2_OK) return st;
            d_kps = (const float2*)d;
        }
    }
    int32_t* d_th = nullptr;
    if ((st = ov2_stage_out(ctx, fast_th_inout, sizeof(int32_t) * (size_t)count, &o, true)) != OV2_OK) return st;
    d_th = (int32_t*)o;
    float2* d_out = nullptr;
    if ((st = ov2_stage_out(ctx, out_pts, sizeof(float) * 2 * (size_t)count * max_per_frame, &o)) != OV2_OK) return st;
    d_out = (float2*)o;
    int32_t* d_cnt = nullptr;
    if ((st = ov2_stage_out(ctx, out_counts, sizeof(int32_t) * (size_t)count, &o)) != OV2_OK) return st;
    d_cnt = (int32_t*)o;
    int2* d_int = nullptr;
    if (out_pts_int) {
        if ((st = ov2_stage_out(ctx, out_pts_int, sizeof(int32_t) * 2 * (size_t)count * max_per_frame, &o)) != OV2_OK) return st;
    } else {
        if ((st = ov2_scratch(ctx, sizeof(int32_t) * 2 * (size_t)count * max_per_frame, &o)) != OV2_OK) return st;
    }
    d_int = (int2*)o;
    uint32_t* d_cand = nullptr;
    int32_t* d_candn = nullptr;
    int32_t* d_ovf = nullptr;
    if ((st = ov2_scratch(ctx, sizeof(uint32_t) * (size_t)count * ncells * cap, &o)) != OV2_OK) return st;
    d_cand = (uint32_t*)o;
    if ((st = ov2_scratch(ctx, sizeof(int32_t) * (size_t)count * ncells, &o)) != OV2_OK) return st;
    d_candn = (int32_t*)o;
    {   // sticky, context-owned flag (reported by this call when it synchronises, else by ov2_batch_end / ov2_frontend_step)
        int* f = nullptr;
        if ((st = ov2_cap_flag_get(ctx, &f)) != OV2_OK) return st;
        d_ovf = (int32_t*)f;
    }

    FastArgs FA;
    FA.img = pyr->l0; FA.w = W; FA.h = H; FA.pitch = (int)pyr->l0_pitch; FA.fstride = (long long)pyr->l0_fstride;
    FA.first = first; FA.cs = cellsize; FA.nwc = nwc; FA.nhc = nhc; FA.cap = cap;
    FA.th = d_th; FA.cand = d_cand; FA.cand_n = d_candn; FA.overflow = d_ovf;
    FA.score_mode = getenv("OV2_FAST_SCORE") ? atoi(getenv("OV2_FAST_SCORE")) : 1;
    if ((st = launch_fast_cells(ctx, pyr, FA, count)) != OV2_OK) return st;
    SweepArgs SA;
    SA.w = W; SA.h = H; SA.cs = cellsize; SA.nwc = nwc; SA.nhc = nhc; SA.cap = cap;
    SA.radius = cellsize / 4; SA.max_per_frame = max_per_frame;
    circle_halfwidths(SA.radius, SA.hw);
    SA.kp_off = d_off; SA.kps = d_kps; SA.cand = d_cand; SA.cand_n = d_candn; SA.th = d_th;
    SA.out_int = d_int; SA.out_n = d_cnt;
    {
        size_t wpr = (size_t)(W + 31) / 32;
        // wavefront sweep (one warp per cell row) unless the bitmap + per-warp sort keys do not fit, cap > 128 or OV2_FAST_SWEEP=seq
        const int nw = nhc < FSWEEP_MAX_WARPS ? nhc : FSWEEP_MAX_WARPS;
        size_t wsmem = ((((size_t)H * wpr + 1) & ~(size_t)1)) * 4 + (size_t)nwc * nhc * 8 + (size_t)nw * cap * 4 + (size_t)nhc * 4 +
                       (size_t)(nhc + 1) * (nwc + 1);
        wsmem = (wsmem + 15) & ~(size_t)15;
        const char* fs = getenv("OV2_FAST_SWEEP");
        if (cap <= 128 && wsmem <= 200 * 1024 && !(fs && fs[0] == 's')) {
            static size_t wave_attr[16] = {0};
            if (wsmem > 48 * 1024 && wsmem > wave_attr[ctx->device & 15]) {
                OV2_CUDA(ctx, cudaFuncSetAttribute(fast_sweep_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                wave_attr[ctx->device & 15] = wsmem;
            }
            OV2_LAUNCH(ctx, "fast_sweep_kernel", fast_sweep_wave_kernel<<<count, nw * 32, wsmem, ctx->stream>>>(SA));
        } else {
        size_t smem = (size_t)H * wpr * 4 + (size_t)cap * 4 + (size_t)(nhc + 1) * (nwc + 1);
        smem = (smem + 15) & ~(size_t)15;
        if (smem > 227 * 1024) return ov2_fail(ctx, OV2_ERR_CAPACITY, "ov2_grid_fast: image too large for the shared-memory mask bitmap");
        if (smem > 48 * 1024)
            OV2_CUDA(ctx, cudaFuncSetAttribute(fast_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OV2_LAUNCH(ctx, "fast_sweep_kernel", fast_sweep_kernel<<<count, 32, smem, ctx->stream>>>(SA));
        }
    }
    SubpixArgs PA;
    PA.img = pyr->l0; PA.w = W; PA.h = H; PA.pitch = (int)pyr->l0_pitch; PA.fstride = (long long)pyr->l0_fstride;
    PA.first = first; PA.max_per_frame = max_per_frame; PA.n = count * max_per_frame;
    PA.in_int = d_int; PA.out = d_out; PA.do_subpix = do_subpix;
    for (int i = 0; i < 7; ++i) {
        float y = (float)(i - 3) / 3;
        float vy = expf(-y * y);
        for (int j = 0; j < 7; ++j) {
            float x = (float)(j - 3) / 3;
            PA.mask[i * 7 + j] = (float)(vy * expf(-x * x));
        }
    }
    OV2_LAUNCH(ctx, "subpix_kernel", subpix_kernel<<<div_up(PA.n, 4), 128, 0, ctx->stream>>>(PA));
    // capacity overflow is an error, never a silent truncation: the flag's D2H mirror is enqueued behind the kernels (it is
    // captured into the composite step's CUDA graph too); a call that synchronises reports it here, a call that only
    // enqueues (batch mode, device outputs) leaves it to the next synchronising call of this context
    const bool host_out = !ctx->pending.empty() && !ctx->batch;
    if ((st = ov2_cap_flag_mirror(ctx)) != OV2_OK) return st;
    st = ov2_end(ctx);
    if (st != OV2_OK) return st;
    if (host_out) return ov2_cap_flag_check(ctx, "ov2_grid_fast: per-cell candidate capacity exceeded");
    return OV2_OK;
}

// Test hook: the per-cell FAST candidate lists of one frame (stage F1 only), so the tests can
// pin cv::FastFeatureDetector parity cell by cell.  cand_out: ncells*cap words
// ((response << 16) | (y << 8) | x, cell-local), cand_n_out: ncells counts; *cap_out = capacity used.
extern "C" ov2_status ov2_debug_fast_cells(ov2_ctx* ctx, const ov2_pyr* pyr, int frame, int cellsize, int fast_th,
                                           uint32_t* cand_out, int32_t* cand_n_out, int cap_in, int* cap_out) {
    if (!ctx || !pyr || !pyr->l0 || frame < 0 || frame >= pyr->batch || cellsize < 8 || cellsize > MAX_CELL)
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_debug_fast_cells: bad arguments");
    const int W = pyr->w[0], H = pyr->h[0];
    const int nwc = W / cellsize, nhc = H / cellsize, ncells = nwc * nhc;
    const int in_w = cellsize - 6;
    int cap = ((in_w + 1) / 2) * ((in_w + 1) / 2);
    if (cap > 128) cap = 128;
    if (cap < 4) cap = 4;
    if (cap_out) *cap_out = cap;
    if (cap_in < cap) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_debug_fast_cells: cap_in too small");
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    void* o = nullptr;
    FastArgs FA;
    FA.img = pyr->l0; FA.w = W; FA.h = H; FA.pitch = (int)pyr->l0_pitch; FA.fstride = (long long)pyr->l0_fstride;
    FA.first = frame; FA.cs = cellsize; FA.nwc = nwc; FA.nhc = nhc; FA.cap = cap;
    if ((st = ov2_scratch(ctx, sizeof(int32_t), &o)) != OV2_OK) return st;
    int32_t* d_th = (int32_t*)o;
    OV2_CUDA(ctx, cudaMemcpyAsync(d_th, &fast_th, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    FA.th = d_th;
    if ((st = ov2_stage_out(ctx, cand_out, sizeof(uint32_t) * (size_t)ncells * cap, &o)) != OV2_OK) return st;
    FA.cand = (uint32_t*)o;
    if ((st = ov2_stage_out(ctx, cand_n_out, sizeof(int32_t) * (size_t)ncells, &o)) != OV2_OK) return st;
    FA.cand_n = (int32_t*)o;
    if ((st = ov2_scratch(ctx, sizeof(int32_t), &o)) != OV2_OK) return st;
    FA.overflow = (int32_t*)o;
    OV2_CUDA(ctx, cudaMemsetAsync(FA.overflow, 0, sizeof(int32_t), ctx->stream));
    FA.score_mode = getenv("OV2_FAST_SCORE") ? atoi(getenv("OV2_FAST_SCORE")) : 1;
    if ((st = launch_fast_cells(ctx, pyr, FA, 1)) != OV2_OK) return st;
    return ov2_end(ctx);
}

// Internal (not part of the C ABI): cornerSubPix stage S on integer points produced by another detector
// (frontend_sscale.cu).  Same launch as inside ov2_grid_fast; slots holding (-1, -1) stay empty.
ov2_status ov2_subpix_launch(ov2_ctx* ctx, const ov2_pyr* pyr, int first, int count, int max_per_frame, const int2* d_int,
                             float2* d_out, int do_subpix) {
    SubpixArgs PA;
    PA.img = pyr->l0; PA.w = pyr->w[0]; PA.h = pyr->h[0]; PA.pitch = (int)pyr->l0_pitch; PA.fstride = (long long)pyr->l0_fstride;
    PA.first = first; PA.max_per_frame = max_per_frame; PA.n = count * max_per_frame;
    PA.in_int = d_int; PA.out = d_out; PA.do_subpix = do_subpix;
    for (int i = 0; i < 7; ++i) {
        float y = (float)(i - 3) / 3;
        float vy = expf(-y * y);
        for (int j = 0; j < 7; ++j) {
            float x = (float)(j - 3) / 3;
            PA.mask[i * 7 + j] = (float)(vy * expf(-x * x));
        }
    }
    OV2_LAUNCH(ctx, "subpix_kernel", subpix_kernel<<<div_up(PA.n, 4), 128, 0, ctx->stream>>>(PA));
    return OV2_OK;
}
