// zignal-b200 :: K3t -- the dense FIR on the 5th-generation tensor cores (tcgen05, TMEM accumulators), FAST mode.
//
// What it replaces: the same graph as kernels/zg_fir.cuh -- c0*_1 + c1*_1[_1] + ... + c(N-1)*_1[_(N-1)], N <= 256,
// which the reference ticks as a 256-term expression over one shifted std::array (flowz/flowz.hpp:130-148, :769-772).
// With the taps shared by all channels that sum IS a dense channel x tap contraction (BASELINE north_star:
// "tensor cores used only where a stage is genuinely a dense channel x tap contraction"):
//
//     Y[c, t] = sum_j X[c, j] * h[t - j]          Y = X * G^T,   G[t, j] = h[t - j]  (banded Toeplitz)
//
// Shape of the MMAs (tcgen05.mma.cta_group::1.kind::tf32, M = 128, K = 8, N = 32 .. 256):
//     D[128 channels, N output times]  +=  A[128 channels, 8 input times]  x  B[N output times, 8 input times]^T
//   * A is a slice of an input block: 128 channel rows x 32 samples (16 KB) exactly as TMA lands it with
//     SWIZZLE_128B -- the canonical K-major operand layout, no repacking;
//   * B is a window of ONE matrix G[r][j] = h[r - j], r = 0..287, j = 0..31 (36 KB, built once per CTA): input block
//     jb (samples 32 jb ..) reaches output times 32 jb .. 32 jb + 287, and the coefficient of input sample j for
//     output time r (both relative to the block) is h[r - j] whatever the block -- the band is the SAME matrix for
//     every block, only the accumulator columns it lands on move;
//   * D: the 512 TMEM columns are a ring of four 128-sample output tiles, zeroed at the start and again by the
//     epilogue as it drains a tile, so every MMA accumulates.  Block jb adds into the 288 columns of output times
//     32 jb .. 32 jb + 287: two MMAs (N = 160 + 128; N <= 256 per instruction) per K-step and operand product, cut once
//     more where the window wraps around the ring.  A tile is complete after the block that ends it and is drained
//     by the epilogue warps while the next three fill.
//
// fp32 accuracy out of TF32 operands (3xTF32): x = hi + lo with hi = cvt.rna.tf32(x), lo = x - hi (exact in fp32),
// the same for the taps; three MMAs hi*hi + lo*hi + hi*lo per K-step, the dropped lo*lo term and the conversion of
// the lo operands are <= 2^-22 |x||h| each.  Accumulation is fp32 in TMEM, in blocked order -- not the reference's
// left-to-right association, hence FAST mode only (bar: 1e-5 block-relative against fir_direct, tests/).
//
// Roles (448 threads): warps 0-3 epilogue (TMEM -> registers -> swizzled staging box -> TMA store; one thread per
// channel row), warps 4-11 split every landed block into hi / lo in shared memory, warp 12 = TMA producer, warp 13 =
// MMA issuer (one lane).  CTAs are persistent: each takes a contiguous range of (channel group, output
// tile) work items; a range that starts in the middle of a row of tiles re-reads its 8 history blocks from the input,
// a range that starts at tile 0 reads them from the delay-line state ([N-1][channels], oldest first).  The state
// after the block is written by a separate small kernel (zg_fir_state_kernel), so nothing here writes state.
#pragma once
#include "zg_stream.cuh"

namespace zgk {

struct FirTcArgs {
    TensorMap in_map;               // planar 2-D {T, C}, box {32 samples, 128 channels}, SWIZZLE_128B
    TensorMap out_map;              // the same over the output block
    const float* state_in;          // [n_taps-1][ch_stride]; slot s = x[t0 - (n_taps-1) + s]
    const float* taps;              // [n_taps]
    long long ch_stride;
    int channels, n_samples, n_taps;
    int n_groups;                   // ceil(channels / 128)
    int n_tiles;                    // ceil(n_samples / 128)
    long long* prof;                // tuning only (ZG_TUNE_FIR_PROF): [grid][8] cycles the roles of a CTA spent waiting
    int split_mode;                 // experiment (ZG_TUNE_FIR_SPLIT): 0 = hi rewritten as tf32(x) (default); 1 = hi left as the
                                    // raw fp32 block, lo = x - trunc_tf32(x); 2 = hi raw, lo = x - rna_tf32(x)
};

constexpr int kTcStages = 4;
constexpr int kTcBlockBytes = 16384;                     // 128 channel rows x 128 bytes
constexpr int kTcGRows = 288;                            // output times one input block reaches (256 taps + 32 samples)
constexpr int kTcGBytes = kTcGRows * 128;                // 36 864
constexpr int kTcSmemG = 2 * kTcGBytes;                  // hi, lo
constexpr int kTcSmemX = 2 * kTcStages * kTcBlockBytes;  // hi[stages], lo[stages]
constexpr int kTcSmemOut = kTcBlockBytes;                // epilogue staging: 128 channels x 32 output times
constexpr int kTcSmemBars = 256;
constexpr int kTcSmemBytes = 1024 /*alignment slack*/ + kTcSmemG + kTcSmemX + kTcSmemOut + kTcSmemBars;
constexpr int kTcSplitWarps = 8;                         // warps 4 .. 4 + kTcSplitWarps - 1
constexpr int kTcThreads = (4 + kTcSplitWarps + 2) * 32;
// instruction descriptor: D = f32, A = B = tf32, both K-major, dense, M = 128; N (a multiple of 16) is added per MMA
constexpr unsigned kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
__device__ __forceinline__ unsigned tc_idesc(unsigned n) { return kTcIdesc | ((n >> 3) << 17); }

// ---- PTX wrappers (tcgen05) ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc_512(unsigned* smem_dst) {          // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_dst)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc_512(unsigned taddr) {            // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {        // arrives when all MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ unsigned long long tc_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) /*LBO, unused with swizzle*/ |
           ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&r)[32]) {   // this warp's 32 lanes x 32 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st32_zero(unsigned taddr) {                // zero this warp's 32 lanes x 32 columns
    const unsigned z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32_rna(float v) {                        // nearest TF32 (10 explicit mantissa bits)
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ int floor_div4(int v) { return v >= 0 ? v >> 2 : -((3 - v) >> 2); }

// The work items of this CTA as runs inside one channel group: f(group, o_begin, o_end)
template <class F>
__device__ __forceinline__ void tc_for_each_run(const FirTcArgs& a, F&& f) {
    const long long total = (long long)a.n_groups * a.n_tiles;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    long long w = (long long)blockIdx.x * per;
    const long long w_end = w + per < total ? w + per : total;
    while (w < w_end) {
        const int g = (int)(w / a.n_tiles), o_s = (int)(w % a.n_tiles);
        const long long left = w_end - w;
        const int o_e = o_s + left < a.n_tiles ? (int)(o_s + left) : a.n_tiles;
        f(g, o_s, o_e);
        w += o_e - o_s;
    }
}

__device__ __forceinline__ void fir_tc_block(const FirTcArgs& a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char* g_hi = base;
    unsigned char* g_lo = base + kTcGBytes;
    unsigned char* x_hi = base + kTcSmemG;
    unsigned char* x_lo = x_hi + kTcStages * kTcBlockBytes;
    unsigned char* y_stage = base + kTcSmemG + kTcSmemX;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + kTcSmemG + kTcSmemX + kTcSmemOut);
    unsigned long long* full = bars;                     // [stages]  TMA landed (or: stage free for a state fill)
    unsigned long long* ready = bars + kTcStages;        // [stages]  hi / lo written
    unsigned long long* empty = bars + 2 * kTcStages;    // [stages]  MMAs that read the stage are done
    unsigned long long* acc_full = bars + 3 * kTcStages; // [4]       accumulator complete
    unsigned long long* acc_empty = acc_full + 4;        // [4]       accumulator drained
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(acc_empty + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.n_taps - 1;                          // depth of the delay line

    // ---- one-time setup: barriers, TMEM, the Toeplitz image G (hi / lo) ----
    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&ready[i], kTcSplitWarps * 32);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_barrier_init();
        prefetch_tmap(&a.in_map);
        prefetch_tmap(&a.out_map);
    }
    if (warp == 0) tc_alloc_512(tmem_slot);
    for (int e = tid; e < kTcGRows * 32; e += kTcThreads) {
        const int r = e >> 5, j = e & 31;
        const int k = r - j;                             // tap index
        const float h = (k >= 0 && k < a.n_taps) ? a.taps[k] : 0.f;
        const float hi = tf32_rna(h);
        const unsigned off = (unsigned)r * 128u + ((((unsigned)j >> 2) ^ ((unsigned)r & 7u)) << 4) + ((unsigned)j & 3u) * 4u;
        *reinterpret_cast<float*>(g_hi + off) = hi;
        *reinterpret_cast<float*>(g_lo + off) = tf32_rna(h - hi);
    }
    fence_proxy_async();                                 // G: generic-proxy writes -> visible to the tensor core
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const unsigned tmem = *tmem_slot;
    // every MMA accumulates: the ring starts zeroed, and the epilogue zeroes a tile again as it drains it
    if (warp < 4) {
        for (int c = 0; c < 512; c += 32) tc_st32_zero(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)c);
        tc_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (warp == 4 + kTcSplitWarps) {
        // ===== TMA producer =====
        if (lane == 0) {
            int n = 0;
            long long w_empty = 0;
            tc_for_each_run(a, [&](int g, int o_s, int o_e) {
                for (int jb = 4 * o_s - 8; jb <= 4 * o_e - 1; ++jb, ++n) {
                    const int st = n % kTcStages;
                    const long long t0 = a.prof ? clock64() : 0;
                    mbar_wait(&empty[st], (unsigned)(((n / kTcStages) & 1) ^ 1));
                    if (a.prof) w_empty += clock64() - t0;
                    if (jb >= 0) {
                        mbar_expect_tx(&full[st], kTcBlockBytes);
                        tma_load_2d(x_hi + st * kTcBlockBytes, &a.in_map, 32 * jb, 128 * g, &full[st]);
                    } else {
                        mbar_arrive(&full[st]);          // before the stream began: the split warps fill it from the state
                    }
                }
            });
            if (a.prof) a.prof[(long long)blockIdx.x * 8 + 5] = w_empty;
        }
    } else if (warp == 5 + kTcSplitWarps) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int n = 0, q_base = 0;
            long long w_full = 0, w_ready = 0, w_acc = 0;
            const long long t_begin = clock64();
            tc_for_each_run(a, [&](int g, int o_s, int o_e) {
                (void)g;
                for (int jb = 4 * o_s - 8; jb <= 4 * o_e - 1; ++jb, ++n) {
                    const int st = n % kTcStages;
                    const unsigned ph = (unsigned)((n / kTcStages) & 1);
                    long long t0 = a.prof ? clock64() : 0;
                    mbar_wait(&full[st], ph);
                    long long t1 = a.prof ? clock64() : 0;
                    mbar_wait(&ready[st], ph);
                    if (a.prof) { w_full += t1 - t0; w_ready += clock64() - t1; }
                    tc_fence_after_sync();
                    // output times this block reaches, clipped to the tiles of this run; the last 32 are new
                    const int t_run0 = 128 * o_s, t_run1 = 128 * o_e;
                    const int t_lo = 32 * jb > t_run0 ? 32 * jb : t_run0;
                    const int t_new = 32 * jb + 256;                    // the last 32 of them are touched for the first time
                    const int t_hi = t_new + 32 < t_run1 ? t_new + 32 : t_run1;
                    const unsigned xa_hi = smem_u32(x_hi + st * kTcBlockBytes), xa_lo = smem_u32(x_lo + st * kTcBlockBytes);
                    const unsigned gb_hi = smem_u32(g_hi), gb_lo = smem_u32(g_lo);
                    if ((jb & 3) == 0 && t_new < t_run1) {       // the new columns open tile jb/4 + 2: its slot must be drained
                        const int q = q_base + (jb / 4 + 2 - o_s);
                        const long long tw = a.prof ? clock64() : 0;
                        mbar_wait(&acc_empty[q & 3], (unsigned)(((q >> 2) & 1) ^ 1));
                        if (a.prof) w_acc += clock64() - tw;
                        tc_fence_after_sync();
                    }
                    // column of output time t in the ring: tile o_s sits in slot q_base & 3
                    const int col0 = (q_base & 3) * 128 - t_run0;
                    // D[:, t_lo .. t_hi) += X_blk * G[t_lo - 32 jb .. t_hi - 32 jb)^T: the 288 output times this block reaches
                    // (fewer at the ends of a run).  An MMA takes N <= 256 columns, so the window is cut once in the middle
                    // (160 + 128: a lone N = 32 MMA costs ~70 cycles for 16 cycles of work) and where the ring wraps.  All
                    // MMAs of a piece are issued back to back: changing the accumulator between MMAs stalls the pipe.
                    struct Seg { int c, n, row; };
                    Seg seg[2];
                    int n_seg = 0;
                    {
                        const int c_lo = (col0 + t_lo) & 511, len = t_hi - t_lo;
                        // one cut: at the end of the ring if the window wraps, else in the middle of a full window
                        const int first = c_lo + len > 512 ? 512 - c_lo : len > 256 ? 160 : len;
                        seg[n_seg++] = Seg{c_lo, first, t_lo - 32 * jb};
                        if (first < len) seg[n_seg++] = Seg{(c_lo + first) & 511, len - first, t_lo + first - 32 * jb};
                    }
                    for (int i = 0; i < n_seg; ++i) {
                        const unsigned idesc = tc_idesc((unsigned)seg[i].n);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const unsigned long long ah = tc_desc(xa_hi + ks * 32), al = tc_desc(xa_lo + ks * 32);
                            const unsigned long long bh = tc_desc(gb_hi + (unsigned)seg[i].row * 128u + ks * 32);
                            const unsigned long long bl = tc_desc(gb_lo + (unsigned)seg[i].row * 128u + ks * 32);
                            tc_mma_tf32(tmem + (unsigned)seg[i].c, ah, bh, idesc, 1u);
                            tc_mma_tf32(tmem + (unsigned)seg[i].c, al, bh, idesc, 1u);
                            tc_mma_tf32(tmem + (unsigned)seg[i].c, ah, bl, idesc, 1u);
                        }
                    }
                    if ((jb & 3) == 3 && floor_div4(jb) >= o_s) { // this block ends tile floor(jb / 4)
                        const int q = q_base + (floor_div4(jb) - o_s);
                        tc_commit(&acc_full[q & 3]);
                    }
                    tc_commit(&empty[st]);
                }
                q_base += o_e - o_s;
            });
            if (a.prof) {
                long long* pr = a.prof + (long long)blockIdx.x * 8;
                pr[0] = clock64() - t_begin; pr[1] = w_full; pr[2] = w_ready; pr[3] = w_acc; pr[4] = n;
            }
        }
    } else if (warp >= 4) {
        // ===== split warps: hi = tf32(x), lo = x - hi, in place / into the lo tile =====
        const int ts = tid - 128;                        // 0 .. 32 * kTcSplitWarps - 1
        int n = 0;
        tc_for_each_run(a, [&](int g, int o_s, int o_e) {
            for (int jb = 4 * o_s - 8; jb <= 4 * o_e - 1; ++jb, ++n) {
                const int st = n % kTcStages;
                mbar_wait(&full[st], (unsigned)((n / kTcStages) & 1));
                unsigned char* hi = x_hi + st * kTcBlockBytes;
                unsigned char* lo = x_lo + st * kTcBlockBytes;
                if (jb >= 0) {
#pragma unroll
                    for (int i = 0; i < 1024 / (kTcSplitWarps * 32); ++i) {
                        const unsigned off = (unsigned)(ts + kTcSplitWarps * 32 * i) << 4;
                        const float4 v = *reinterpret_cast<const float4*>(hi + off);
                        float4 h, l;
                        if (a.split_mode == 1) {
                            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                        } else {
                            h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                        }
                        // the tensor core truncates an fp32 operand to TF32 (measured: tools/fir_split_check.py), which would
                        // bias every lo term the same way: round it to nearest here instead
                        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                        if (a.split_mode == 0) *reinterpret_cast<float4*>(hi + off) = h;
                        *reinterpret_cast<float4*>(lo + off) = l;
                    }
                } else {
                    // delay-line state: sample time t = 32*jb + j (negative) is slot D + t; older than the line: zero
                    const int ch = g * 128 + (ts & 127);
                    const unsigned r = (unsigned)(ts & 127);
                    for (int j = ts >> 7; j < 32; j += kTcSplitWarps / 4) {
                        const int s = D + 32 * jb + j;
                        const float v = (s >= 0 && ch < a.channels) ? a.state_in[(long long)s * a.ch_stride + ch] : 0.f;
                        const float h = tf32_rna(v);
                        const unsigned off = r * 128u + ((((unsigned)j >> 2) ^ (r & 7u)) << 4) + ((unsigned)j & 3u) * 4u;
                        *reinterpret_cast<float*>(hi + off) = h;
                        *reinterpret_cast<float*>(lo + off) = tf32_rna(v - h);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&ready[st]);
            }
        });
    } else {
        // ===== epilogue warps 0-3: TMEM lanes 32*warp .. +31 = channel rows; columns = output times.  32 columns at a
        //       time go through a staging box in the layout of the input blocks (row = channel, SWIZZLE_128B: the eight
        //       STS.128 of a quarter-warp cover all banks) and leave with one TMA store, which also clips the ragged
        //       end of the block and of the channels =====
        int q = 0;
        long long w_accfull = 0, t_store = 0;
        const unsigned r = (unsigned)(warp * 32 + lane);
        tc_for_each_run(a, [&](int g, int o_s, int o_e) {
            for (int o = o_s; o < o_e; ++o, ++q) {
                const int slot = q & 3;
                const long long t0c = a.prof ? clock64() : 0;
                mbar_wait(&acc_full[slot], (unsigned)((q >> 2) & 1));
                const long long t1c = a.prof ? clock64() : 0;
                tc_fence_after_sync();
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    unsigned v[32];
                    tc_ld32(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(slot * 128 + cc * 32), v);
                    tc_st32_zero(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(slot * 128 + cc * 32));   // ready for the next tile
                    if (tid == 0) tma_wait_read<0>();            // the previous store has read the staging box
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<uint4*>(y_stage + r * 128u + ((((unsigned)i) ^ (r & 7u)) << 4)) =
                            make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tid == 0 && o * 128 + cc * 32 < a.n_samples) {
                        tma_store_2d(&a.out_map, o * 128 + cc * 32, g * 128, y_stage);
                        tma_commit();
                    }
                }
                tc_wait_st();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
                if (a.prof) { w_accfull += t1c - t0c; t_store += clock64() - t1c; }
            }
        });
        if (tid == 0) tma_wait_all<0>();                         // shared memory must outlive the last store
        if (a.prof && tid == 0) { a.prof[(long long)blockIdx.x * 8 + 6] = w_accfull; a.prof[(long long)blockIdx.x * 8 + 7] = t_store; }
    }

    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) tc_dealloc_512(tmem);
}

// The delay line after a block of T samples, oldest first: slot s = x[T - D + s], taken from the input where
// T - D + s >= 0 and from the old line (shifted by T) otherwise.  Separate from the evaluation: the persistent CTAs
// above cut the block anywhere, and the line is ping-ponged between two buffers like K3's.
__global__ void zg_fir_state_kernel(const float* __restrict__ in, long long ld_in, const float* __restrict__ state_in,
                                    float* __restrict__ state_out, long long ch_stride, int channels, int n_samples, int depth) {
    // a 32-channel x 32-slot tile through shared memory: the input is read along time (coalesced per channel row), the
    // line is written along channels (coalesced per slot)
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {                               // r: channel of the tile, tx: slot
        const int ch = c0 + r, s = s0 + tx;
        float v = 0.f;
        if (ch < channels && s < depth) {
            const long long t = (long long)n_samples - depth + s;
            v = t >= 0 ? in[(long long)ch * ld_in + t] : state_in[(long long)(s + n_samples) * ch_stride + ch];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {                               // r: slot of the tile, tx: channel
        const int ch = c0 + tx, s = s0 + r;
        if (ch < channels && s < depth) state_out[(long long)s * ch_stride + ch] = tile[tx][r];
    }
}

}  // namespace zgk
