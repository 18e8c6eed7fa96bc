// zignal-b200 :: K3 -- dense FIR  y[t] = ((c0*x[t] + c1*x[t-1]) + c2*x[t-2]) + ... + c(N-1)*x[t-N+1]
// for many independent channels (BASELINE configs[3]: 256 taps x 32 768 channels), CUDA cores.
//
// What it replaces: the reference ticks this graph as a 256-term expression template over one
// std::array<float,255> delay line that is shifted by one element per sample
// (flowz/flowz.hpp:130-148 rotate_push_back, :950-958 place_delay, :769-772 leaf arithmetic).
// Here nothing is shifted: the delay line of a channel IS its input row, kept in a shared-memory
// ring of TMA boxes, and the N-1 floats of state are only the hand-over between two blocks.
//
// Design (B200):
//   * a CTA owns 32 channels (lane = channel) and a time segment; its W warps compute W consecutive
//     32-sample output boxes per step, all reading the same ring of input boxes (32 channels x 32
//     samples, 4 KB, SWIZZLE_128B: a lane reads its own row 16 bytes at a time, conflict free).
//     The ring holds H = ceil((N-1)/32) history boxes + the W boxes of this step + the W boxes of the
//     next step already in flight (one mbarrier per step parity, lane 0 of warp 0 issues the TMA).
//   * a warp computes a box as two tiles of 16 outputs; per 16 taps it holds the two aligned 16-sample
//     input blocks the tile touches in registers (32 floats) and does 256 mul + 256 add (EXACT) or
//     256 FMA (FAST) per 4 LDS.128 of samples + 4 LDS.128 of taps: the kernel is bound by FP32 issue
//     (511 instructions per output sample in EXACT mode), not by HBM -- see DESIGN.md.
//   * products are added in tap order, separately rounded in EXACT mode (the accumulator starts at
//     -0.0f, and -0.0f + p == p for every p), so the result is bit-identical to the reference's parse
//     tree; FAST mode contracts each step to one FMA.
//   * time is split into segments when there are too few channel groups to fill 148 SMs: a FIR has no
//     recurrence, a later segment just loads its history boxes from the input instead of the state.
//     State is ping-ponged (read state_in, write state_out) because the segment that writes the new
//     state may run before the segment that reads the old one.
//   * outputs go to a per-warp double-buffered box and leave with a TMA store.
//   * interleaved frames ([T][C] buffers): the same ring, but a box is 32 frames of 32 channels (no swizzle), a
//     lane walks down its column with LDS.32 / STS.32 (all 32 lanes on one 128-byte row: conflict free) -- four
//     times the shared-memory instructions of the planar form for the same arithmetic.
#pragma once
#include "zg_stream.cuh"

namespace zgk {

struct FirArgs {
    TensorMap in_map;               // planar: 2-D {T, C}, box 32 x 32, SWIZZLE_128B; interleaved: 2-D {C, T}, no swizzle
    TensorMap out_map;
    const float* state_in;          // [n_taps-1][ch_stride]; slot j = x[t0 - (n_taps-1) + j]
    float* state_out;
    const float* taps;              // [n_taps] (device)
    long long ch_stride;
    int channels;
    int n_samples;
    int n_taps;
    int hist_boxes;                 // H
    int ring_boxes;                 // NR >= H + 2W
    int seg_boxes;                  // boxes per time segment: multiple of W, >= H
    int n_segs;
};

struct FirCursor {                  // a 16-sample block of the ring: (slot, half); moves back in time
    int slot, half;
};

template <bool kInterleaved>
__device__ __forceinline__ void fir_load16(float (&x)[16], const unsigned char* ring, FirCursor c, unsigned row,
                                           unsigned sw) {
    if (kInterleaved) {                                // frame f of the box at f * 128, this lane's channel at `row`
        const unsigned char* base = ring + (size_t)c.slot * kTileBytes + (unsigned)c.half * 2048u + row;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = *reinterpret_cast<const float*>(base + i * 128);
        return;
    }
    const unsigned char* base = ring + (size_t)c.slot * kTileBytes + row;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(base + ((((unsigned)(c.half * 4 + i)) ^ sw) << 4));
        x[4 * i] = v.x;
        x[4 * i + 1] = v.y;
        x[4 * i + 2] = v.z;
        x[4 * i + 3] = v.w;
    }
}

// 16 taps (cs[0..15], delays k0..k0+15) into 16 outputs; xn = the aligned block that holds x[t0-k0 ..],
// xo = the block before it.  kmax < 16 only for the last, partial group of taps.
template <bool kExact, bool kTail>
__device__ __forceinline__ void fir_body(float (&acc)[16], const float (&xn)[16], const float (&xo)[16],
                                         const float* cs, int kmax) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(cs + 4 * i);
        c[4 * i] = v.x;
        c[4 * i + 1] = v.y;
        c[4 * i + 2] = v.z;
        c[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        if (kTail && kk >= kmax) break;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float x = (r - kk >= 0) ? xn[(r - kk) & 15] : xo[(16 + r - kk) & 15];
            if (kExact) acc[r] = __fadd_rn(acc[r], __fmul_rn(c[kk], x));
            else acc[r] = fmaf(c[kk], x, acc[r]);
        }
    }
}

template <bool kExact, bool kInterleaved = false>
__device__ __forceinline__ void fir_block(const FirArgs& a) {
    extern __shared__ __align__(1024) unsigned char smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int W = blockDim.x >> 5;
    const int N = a.n_taps;
    const int H = a.hist_boxes;
    const int NR = a.ring_boxes;
    const int n_groups = (a.channels + 31) / 32;
    const int group = blockIdx.x % n_groups;
    const int seg = blockIdx.x / n_groups;
    const int c0 = group * 32;
    const int ch = c0 + lane;
    const bool ch_ok = ch < a.channels;

    const int total_boxes = (a.n_samples + kTileT - 1) / kTileT;
    const int b_begin = seg * a.seg_boxes;
    const int b_end = total_boxes < b_begin + a.seg_boxes ? total_boxes : b_begin + a.seg_boxes;
    if (b_begin >= b_end) return;                                   // CTA-uniform
    const int n_steps = (b_end - b_begin + W - 1) / W;

    unsigned char* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char* obuf = ring + (size_t)NR * kTileBytes + (size_t)warp * 2 * kTileBytes;
    float* taps_s = reinterpret_cast<float*>(ring + (size_t)NR * kTileBytes + (size_t)W * 2 * kTileBytes);
    const int n_taps_pad = (N + 15) & ~15;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(taps_s + n_taps_pad);   // [0],[1] steps, [2] history

    // local box index lb = b - (b_begin - H) >= 0; ring slot = lb % NR
    auto issue_step = [&](int s) {                                  // thread 0 only
        const int b0 = b_begin + s * W;
        const int nb = b_end - b0 < W ? b_end - b0 : W;
        mbar_expect_tx(&bars[s & 1], (unsigned)nb * kTileBytes);
        for (int i = 0; i < nb; ++i) {
            const int lb = b0 + i - b_begin + H;
            const int tx = kInterleaved ? c0 : (b0 + i) * kTileT, ty = kInterleaved ? (b0 + i) * kTileT : c0;
            tma_load_2d(ring + (size_t)(lb % NR) * kTileBytes, &a.in_map, tx, ty, &bars[s & 1]);
        }
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
        prefetch_tmap(&a.in_map);
        prefetch_tmap(&a.out_map);
        if (seg > 0) {                                              // history = the input before the segment
            mbar_expect_tx(&bars[2], (unsigned)H * kTileBytes);
            for (int i = 0; i < H; ++i) {
                const int t = (b_begin - H + i) * kTileT;
                tma_load_2d(ring + (size_t)i * kTileBytes, &a.in_map, kInterleaved ? c0 : t, kInterleaved ? t : c0, &bars[2]);
            }
        }
        issue_step(0);
    }
    for (int i = tid; i < n_taps_pad; i += blockDim.x) taps_s[i] = i < N ? a.taps[i] : 0.f;
    // byte offset of sample `s` (0..31) of this lane's channel inside a box
    auto elem = [&](unsigned s) {
        return kInterleaved ? s * 128u + (unsigned)lane * 4u
                            : (unsigned)lane * 128u + (((s >> 2) ^ ((unsigned)lane & 7u)) << 4) + (s & 3u) * 4u;
    };
    if (seg == 0) {
        // history = the delay line of the previous block: slot j = x[-(N-1) + j]  ->  local sample
        // 32H - (N-1) + j; older positions are never read by the arithmetic, zero them
        const int first = 32 * H - (N - 1);
#pragma unroll 4
        for (int ls = warp; ls < 32 * H; ls += W) {
            const int j = ls - first;
            const float v = (j >= 0 && ch_ok) ? a.state_in[(long long)j * a.ch_stride + ch] : 0.f;
            const unsigned off = (unsigned)(ls >> 5) * kTileBytes + elem((unsigned)ls & 31u);
            *reinterpret_cast<float*>(ring + off) = v;
        }
        fence_proxy_async();                                        // these slots are overwritten by TMA later
    }
    __syncthreads();
    if (seg > 0) mbar_wait(&bars[2], 0);

    const unsigned row = kInterleaved ? (unsigned)lane * 4u : (unsigned)lane * 128u;
    const unsigned sw = (unsigned)lane & 7u;
    const int n_full = N >> 4;
    const int rem = N & 15;

    for (int s = 0; s < n_steps; ++s) {
        // every warp has left step s-1 (barrier below): the slots of step s+1 (last read as history by
        // step s-1) are free
        if (tid == 0 && s + 1 < n_steps) issue_step(s + 1);
        mbar_wait(&bars[s & 1], (unsigned)((s >> 1) & 1));

        const int b = b_begin + s * W + warp;
        if (b < b_end) {
            unsigned char* ob = obuf + (size_t)(s & 1) * kTileBytes;
            if (lane == 0) tma_wait_read<1>();                      // the store of step s-2 has read `ob`
            __syncwarp();
            const int lb = b - b_begin + H;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                float acc[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[r] = -0.f;
                FirCursor cur{lb % NR, h};
                int back = 2 * lb + h;                              // 16-sample blocks available before `cur`
                auto prev = [&]() {
                    if (cur.half) cur.half = 0;
                    else { cur.half = 1; cur.slot = cur.slot ? cur.slot - 1 : NR - 1; }
                    --back;
                };
                float xa[16], xb[16];
                fir_load16<kInterleaved>(xa, ring, cur, row, sw);
                prev();
                fir_load16<kInterleaved>(xb, ring, cur, row, sw);
                int j = 0;
                for (; j + 2 <= n_full; j += 2) {
                    fir_body<kExact, false>(acc, xa, xb, taps_s + 16 * j, 16);
                    prev();
                    if (back >= 0) fir_load16<kInterleaved>(xa, ring, cur, row, sw);
                    fir_body<kExact, false>(acc, xb, xa, taps_s + 16 * (j + 1), 16);
                    prev();
                    if (back >= 0) fir_load16<kInterleaved>(xb, ring, cur, row, sw);
                }
                if (j < n_full) {
                    fir_body<kExact, false>(acc, xa, xb, taps_s + 16 * j, 16);
                    if (rem) {
                        prev();
                        if (back >= 0) fir_load16<kInterleaved>(xa, ring, cur, row, sw);
                        fir_body<kExact, true>(acc, xb, xa, taps_s + 16 * (j + 1), rem);
                    }
                } else if (rem) {
                    fir_body<kExact, true>(acc, xa, xb, taps_s + 16 * j, rem);
                }
                if (kInterleaved) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) *reinterpret_cast<float*>(ob + (unsigned)(16 * h + r) * 128u + row) = acc[r];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(ob + row + ((((unsigned)(h * 4 + i)) ^ sw) << 4)) =
                            make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&a.out_map, kInterleaved ? c0 : b * kTileT, kInterleaved ? b * kTileT : c0, ob);
                tma_commit();
            }
        }
        __syncthreads();
    }

    // ---- new state: the last N-1 samples of (old delay line ++ block), still in the ring ----
    if (b_end == total_boxes) {
        for (int j = warp; j < N - 1; j += W) {
            const long long t = (long long)a.n_samples - (N - 1) + j;            // block time, may be < 0
            const int ls = (int)(t - 32ll * (b_begin - H));
            const unsigned off = (unsigned)((ls >> 5) % NR) * kTileBytes + elem((unsigned)ls & 31u);
            const float v = *reinterpret_cast<const float*>(ring + off);
            if (ch_ok) a.state_out[(long long)j * a.ch_stride + ch] = v;
        }
    }
    if (lane == 0) tma_wait_all<0>();                               // smem must outlive the last stores
}

}  // namespace zgk
