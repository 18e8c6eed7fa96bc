// zignal-b200 :: section-parallel kernel for cascades of direct-form-1 biquads (K1b).
//
// Same graph, same arithmetic and same state layout as zg_biquad.cuh (reference spelling
// test/benchmark.cpp:25-33, `fwd |= bwd` chained S times), for LOW channel counts.  With one lane per
// channel a 4096-channel block is 128 warps -- less than one warp per SM on a B200, all four
// schedulers of an SM waiting on one serial recurrence.  Here a channel is evaluated by S lanes, lane
// `sec` owning section `sec` of the cascade:
//
//     lane = channel_in_warp * S + sec            32/S channels per warp, S x as many warps
//
// The cascade becomes a systolic pipeline in time, advancing in GROUPS of four samples: in group g
// lane `sec` evaluates samples 4(g-sec) .. 4(g-sec)+3.  Its inputs are the four outputs lane sec-1
// produced one group earlier, handed over with __shfl_up; a whole group (>= 4 x 12 cycles of the
// lane's own recurrence) lies between producing a value and consuming it, so the shuffle latency
// (~25 cycles + the feed-forward half of the section) never sits on the critical path -- with a skew
// of one or two samples it does, and the warp runs at half speed (measured).  Every lane performs
// exactly the operations of its section in exactly the reference's association, so EXACT mode stays
// bit-identical to the oracle -- only the *order in time* in which independent sections are evaluated
// changes, which no result depends on.
//
// A tile (NB boxes of [32/S channels x 32 samples], 128-byte rows, SWIZZLE_128B) is self-contained:
// the pipeline fills at its start and drains at its end (S-1 extra groups per tile, ~2 % at 512
// samples), so between tiles the registers hold the plain delay-line state and the state rows in HBM
// are interchangeable with the lane-per-channel kernel's.  Groups in which some lane is outside the
// tile run predicated ("slow" groups); the steady state in between is one LDS.128 (lane 0 of a
// channel reads chunk g), four unpredicated ticks, four shuffles and one STS.128 (lane S-1 overwrites
// chunk g-(S-1) in place -- read S-1 groups earlier, so no hazard).
#pragma once
#include <type_traits>

#include "zg_stream.cuh"

namespace zgk {

template <bool kExact>
struct Df1Lane {
    float b0, b1, b2, a1, a2;
    float x1, x2, y1, y2;
    __device__ __forceinline__ float eval(float in) const {
        if (kExact) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(b0, in), __fmul_rn(b1, x1)), __fmul_rn(b2, x2));
            return __fadd_rn(__fadd_rn(v, __fmul_rn(a1, y1)), __fmul_rn(a2, y2));
        } else {
            const float v = fmaf(b2, x2, fmaf(b1, x1, b0 * in));
            return fmaf(a2, y2, fmaf(a1, y1, v));
        }
    }
    __device__ __forceinline__ void push(float in, float y) {
        x2 = x1; x1 = in; y2 = y1; y1 = y;
    }
};

template <int S, bool kExact, bool kUniform>
__device__ __forceinline__ void biquad_lanes_block(const StreamArgs& a) {
    static_assert(S == 2 || S == 4, "lanes per channel: 2 or 4 (the box must span whole swizzle atoms)");
    constexpr int CPW = 32 / S;              // channels per warp
    constexpr int kBoxBytes = CPW * 128;

    extern __shared__ __align__(1024) unsigned char smem[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int St = a.stages;
    const int NB = a.boxes;
    const int tile_t = NB * kTileT;
    const int sec = lane & (S - 1);
    const int cl = lane / S;
    const long long gw = (long long)blockIdx.x * warps_per_cta + warp;
    const long long c0ll = gw * CPW;
    if (c0ll >= a.channels) return;                    // warp-uniform
    const int c0 = (int)c0ll;
    const int ch = c0 + cl;
    const bool ch_ok = ch < a.channels;
    const bool first = sec == 0, last = sec == S - 1;

    unsigned char* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const unsigned stage_bytes = (unsigned)NB * kBoxBytes;
    unsigned char* my = tiles + (size_t)warp * St * stage_bytes;
    unsigned long long* bars =
        reinterpret_cast<unsigned long long*>(tiles + (size_t)warps_per_cta * St * stage_bytes) + warp * St;

    // ---- this lane's section: coefficients and delay-line state ----
    // kernel slots as in zg_biquad.cuh: state 2k / 2k+1 = signal k two / one tick ago, params 5k..5k+4
    Df1Lane<kExact> f;
    {
        float c[5];
#pragma unroll
        for (int j = 0; j < 5; ++j)
            c[j] = kUniform ? a.uparams[5 * sec + j] : (ch_ok ? a.params[(long long)(5 * sec + j) * a.ch_stride + ch] : 0.f);
        f.b0 = c[0]; f.b1 = c[1]; f.b2 = c[2]; f.a1 = c[3]; f.a2 = c[4];
        auto ld = [&](int slot) { return ch_ok ? a.state[(long long)a.state_row[slot] * a.ch_stride + ch] : 0.f; };
        f.x2 = ld(2 * sec); f.x1 = ld(2 * sec + 1);
        f.y2 = ld(2 * sec + 2); f.y1 = ld(2 * sec + 3);
    }

    if (lane == 0) {
        for (int i = 0; i < St; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
        prefetch_tmap(&a.in_map[0]);
        prefetch_tmap(&a.out_map[0]);
    }
    __syncwarp();

    const int n_tiles = (a.n_samples + tile_t - 1) / tile_t;
    auto boxes_in_tile = [&](int t0) {
        const int left = (a.n_samples - t0 + kTileT - 1) / kTileT;
        return left < NB ? left : NB;
    };
    auto issue_load = [&](int i) {                     // lane 0 only
        const int slot = i % St;
        const int t0 = i * tile_t;
        const int nb = boxes_in_tile(t0);
        mbar_expect_tx(&bars[slot], (unsigned)(nb * kBoxBytes));
        unsigned char* dst = my + (size_t)slot * stage_bytes;
        for (int b = 0; b < nb; ++b) tma_load_2d(dst + b * kBoxBytes, &a.in_map[0], t0 + b * kTileT, c0, &bars[slot]);
    };
    if (lane == 0) {
        const int pre = n_tiles < St - 1 ? n_tiles : St - 1;
        for (int i = 0; i < pre; ++i) issue_load(i);
    }

    // row `cl` of every box; 16-byte chunk j of the row lives at chunk j ^ (cl & 7)   (SWIZZLE_128B).
    // The eight chunk offsets of this lane never change: keep them in registers so that the steady
    // state addresses shared memory as [uniform box base + off[j]] with no per-access arithmetic.
    unsigned off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = (unsigned)cl * 128u + ((((unsigned)j) ^ (unsigned)(cl & 7)) << 4);

    for (int i = 0; i < n_tiles; ++i) {
        const int slot = i % St;
        const int t0 = i * tile_t;
        const int nb = boxes_in_tile(t0);
        const int nt = a.n_samples - t0 < tile_t ? a.n_samples - t0 : tile_t;
        unsigned char* stage = my + (size_t)slot * stage_bytes;
        auto sample_ptr = [&](int t) {                 // slow groups only
            const unsigned g = (unsigned)t >> 2;
            return reinterpret_cast<float*>(stage + (g >> 3) * kBoxBytes + (unsigned)cl * 128u +
                                            (((g & 7u) ^ (unsigned)(cl & 7)) << 4) + ((unsigned)t & 3u) * 4u);
        };

        mbar_wait(&bars[slot], (unsigned)((i / St) & 1));

        float r[4] = {0.f, 0.f, 0.f, 0.f};             // outputs of lane sec-1 for the samples of this group
        auto slow_group = [&](int g) {                 // predicated: lanes may be outside the tile
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = 4 * (g - sec) + q;
                const bool act = ch_ok && m >= 0 && m < nt;
                float in = r[q];
                if (first && act) in = *sample_ptr(m);
                const float y = f.eval(in);
                if (act) f.push(in, y);
                if (last && act) *sample_ptr(m) = y;
                o[q] = y;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = __shfl_up_sync(0xffffffffu, o[q], 1);
        };
        // steady state: chunk j of the box at `box` in, chunk j-(S-1) (possibly of the box before) out
        // (the input chunk was loaded one group ahead: a warp alone on its scheduler cannot hide LDS latency)
        float4 xcur = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fast_group = [&](unsigned char* box, auto jc, bool more) {
            constexpr int j = decltype(jc)::value;
            constexpr int jo = (j - (S - 1) + 8) & 7;
            constexpr int back = (j - (S - 1)) < 0 ? kBoxBytes : 0;
            const float4 xv = xcur;
            if (j < 7) xcur = *reinterpret_cast<const float4*>(box + off[(j + 1) & 7]);
            else if (more) xcur = *reinterpret_cast<const float4*>(box + kBoxBytes + off[0]);
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float xq = q == 0 ? xv.x : q == 1 ? xv.y : q == 2 ? xv.z : xv.w;
                const float in = first ? xq : r[q];
                o[q] = f.eval(in);
                f.push(in, o[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = __shfl_up_sync(0xffffffffu, o[q], 1);
            if (last) *reinterpret_cast<float4*>(box - back + off[jo]) = make_float4(o[0], o[1], o[2], o[3]);
        };

        const int nfull = nt >> 5;                     // boxes without a ragged tail
        const int total = (nt + 3) / 4 + (S - 1);      // groups until the last lane has drained
        if (nfull >= 1) {
            for (int g = 0; g < S - 1; ++g) slow_group(g);
            // rest of box 0: every lane is inside the tile from group S-1 on
            xcur = *reinterpret_cast<const float4*>(stage + off[S - 1]);
            if constexpr (S - 1 <= 1) fast_group(stage, std::integral_constant<int, 1>{}, true);
            if constexpr (S - 1 <= 2) fast_group(stage, std::integral_constant<int, 2>{}, true);
            fast_group(stage, std::integral_constant<int, 3>{}, true);
            fast_group(stage, std::integral_constant<int, 4>{}, true);
            fast_group(stage, std::integral_constant<int, 5>{}, true);
            fast_group(stage, std::integral_constant<int, 6>{}, true);
            fast_group(stage, std::integral_constant<int, 7>{}, nfull > 1);
#pragma unroll 1
            for (int b = 1; b < nfull; ++b) {
                unsigned char* box = stage + (unsigned)b * kBoxBytes;
                fast_group(box, std::integral_constant<int, 0>{}, true);
                fast_group(box, std::integral_constant<int, 1>{}, true);
                fast_group(box, std::integral_constant<int, 2>{}, true);
                fast_group(box, std::integral_constant<int, 3>{}, true);
                fast_group(box, std::integral_constant<int, 4>{}, true);
                fast_group(box, std::integral_constant<int, 5>{}, true);
                fast_group(box, std::integral_constant<int, 6>{}, true);
                fast_group(box, std::integral_constant<int, 7>{}, b + 1 < nfull);
            }
            for (int g = nfull * 8; g < total; ++g) slow_group(g);
        } else {
            for (int g = 0; g < total; ++g) slow_group(g);
        }

        fence_proxy_async();                           // generic-proxy writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
            for (int b = 0; b < nb; ++b) tma_store_2d(&a.out_map[0], t0 + b * kTileT, c0, stage + b * kBoxBytes);
            tma_commit();
            const int nxt = i + St - 1;
            if (nxt < n_tiles) {
                tma_wait_read<1>();                    // the slot of tile i-1: its store has left smem
                issue_load(nxt);
            }
        }
    }

    if (lane == 0) tma_wait_all<0>();

    // ---- state back to HBM (pipeline drained: plain delay lines) ----
    if (ch_ok) {
        auto st = [&](int slot, float v) { a.state[(long long)a.state_row[slot] * a.ch_stride + ch] = v; };
        st(2 * sec + 2, f.y2);
        st(2 * sec + 3, f.y1);
        if (first) { st(0, f.x2); st(1, f.x1); }
    }
}

}  // namespace zgk
