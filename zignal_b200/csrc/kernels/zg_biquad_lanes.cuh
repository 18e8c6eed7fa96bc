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
// The cascade becomes a systolic pipeline in time, advancing in GROUPS of four samples: in iteration
// g lane `sec` evaluates samples 4c .. 4c+3 of chunk c = g - 2*sec.  Values travel between
// neighbouring lanes through a 512-byte per-warp exchange buffer in shared memory: at the end of an
// iteration every lane stores its four outputs with ONE STS.128 -- the last lane of a channel into the
// output tile, the others into their exchange slot -- and at the start of the next iteration every
// lane issues ONE LDS.128 -- the first lane of a channel from the input tile, the others from the
// slot of the lane below -- whose result it consumes one iteration later.  Two iterations (>= 100
// cycles) therefore lie between producing a value and needing it, so neither the shared-memory
// round trip nor the feed-forward half of the next section is ever on the critical path, which is the
// lane's own 12-cycle recurrence (a shuffle-based hand-over with a skew of 1-4 samples measured 2-3x
// slower: its latency lands on the path all lanes of the warp share).  Every lane performs exactly
// the operations of its section in exactly the reference's association, so EXACT mode stays
// bit-identical to the oracle -- only the *order in time* in which independent sections are evaluated
// changes, which no result depends on.
//
// A tile (NB boxes of [32/S channels x 32 samples], 128-byte rows, SWIZZLE_128B; one 3-D TMA
// operation per tile and direction) is self-contained: the pipeline fills at its start and drains at
// its end (2(S-1) extra iterations per tile, < 5 % at 512 samples), so between tiles the registers
// hold the plain delay-line state and the state rows in HBM are interchangeable with the
// lane-per-channel kernel's.  Iterations in which some lane is outside the tile run predicated
// ("slow"); the steady state in between is one LDS.128, four unpredicated ticks and one STS.128.
#pragma once
#include <type_traits>

#include "zg_stream.cuh"

namespace zgk {

__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* map, int x, int y, int z,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z),
        "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const TensorMap* map, int x, int y, int z, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                 "r"(x), "r"(y), "r"(z), "r"(smem_u32(src))
                 : "memory");
}

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
    // push only where `act`; written with selp so that the fill / drain iterations stay branch-free
    // (as `if (act) push(..)` they compiled to one divergent branch per sample: 5x the steady state)
    static __device__ __forceinline__ float sel(bool p, float a, float b) {
        float r;
        asm("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\nselp.f32 %0, %1, %2, q;\n}" : "=f"(r) : "f"(a), "f"(b), "r"((unsigned)p));
        return r;
    }
    __device__ __forceinline__ void push_if(bool act, float in, float y) {
        x2 = sel(act, x1, x2); x1 = sel(act, in, x1); y2 = sel(act, y1, y2); y1 = sel(act, y, y1);
    }
};

constexpr int kLanesXbufBytes = 1024;        // exchange buffer per warp: 2 chunks x 32 lanes x 16 bytes

// in_map[0] / out_map[0]: 2-D {T, C}, box {32, 32/S}        (ragged last tile, box by box)
// in_map[1] / out_map[1]: 3-D {32, C, T/32} over the full 32-sample boxes, box {32, 32/S, NB}
template <int S, bool kExact, bool kUniform, int CH = 1>
__device__ __forceinline__ void biquad_lanes_block(const StreamArgs& a) {
    static_assert(CH == 1 || CH == 2, "16-byte chunks (of four samples) per iteration");
    static_assert(S == 2 || S == 4, "lanes per channel: 2 or 4 (the box must span whole swizzle atoms)");
    constexpr int CPW = 32 / S;              // channels per warp
    constexpr int kBoxBytes = CPW * 128;
    constexpr int LAG = 2;                   // iterations between neighbouring sections
    constexpr int DRAIN = LAG * (S - 1);     // iterations until the last section has caught up
    constexpr int IPB = 8 / CH;              // iterations per 32-sample box

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
    unsigned char* xbuf = tiles + (size_t)warps_per_cta * St * stage_bytes + (size_t)warp * kLanesXbufBytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
                                   tiles + (size_t)warps_per_cta * (St * stage_bytes + kLanesXbufBytes)) + warp * St;
    // exchange buffer: CH arrays of 32 x 16 bytes (array u = chunk u of the iteration), so that the 32
    // lanes of one STS.128 / LDS.128 touch 512 contiguous bytes
    unsigned char* const x_in = xbuf + (lane > 0 ? lane - 1 : 0) * 16;   // what the lane below stored
    unsigned char* const x_out = xbuf + lane * 16;

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
        if (a.flags & 1) {
            prefetch_tmap(&a.in_map[1]);
            prefetch_tmap(&a.out_map[1]);
        }
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) *reinterpret_cast<float4*>(x_out + u * 512) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const int n_tiles = (a.n_samples + tile_t - 1) / tile_t;
    const int full_boxes = a.n_samples >> 5;           // boxes of the block without a ragged tail
    const bool ragged = (a.n_samples & 31) != 0;
    const bool have3d = (a.flags & 1) != 0;
    // tile i: boxes [i*NB, i*NB + nbf) are full; `rag` = the block's ragged box follows them in this tile
    auto tile_shape = [&](int i, int& nbf, bool& rag) {
        const int b0 = i * NB;
        nbf = full_boxes - b0;
        nbf = nbf < 0 ? 0 : (nbf > NB ? NB : nbf);
        rag = ragged && full_boxes >= b0 && full_boxes < b0 + NB;
        if (!have3d) {                                 // no whole-tile map: every box goes on its own
            if (!rag) nbf -= 1;
            rag = true;
        }
    };
    auto issue_load = [&](int i) {                     // lane 0 only
        const int slot = i % St;
        int nbf; bool rag;
        tile_shape(i, nbf, rag);
        unsigned char* dst = my + (size_t)slot * stage_bytes;
        if (!rag) {                                    // one operation; boxes past the end are zero-filled
            mbar_expect_tx(&bars[slot], (unsigned)(NB * kBoxBytes));
            tma_load_3d(dst, &a.in_map[1], 0, c0, i * NB, &bars[slot]);
        } else {
            mbar_expect_tx(&bars[slot], (unsigned)((nbf + 1) * kBoxBytes));
            for (int b = 0; b <= nbf; ++b)
                tma_load_2d(dst + b * kBoxBytes, &a.in_map[0], (i * NB + b) * kTileT, c0, &bars[slot]);
        }
    };
    if (lane == 0) {
        const int pre = n_tiles < St - 1 ? n_tiles : St - 1;
        for (int i = 0; i < pre; ++i) issue_load(i);
    }

    // row `cl` of every box; 16-byte chunk j of the row lives at chunk j ^ (cl & 7)   (SWIZZLE_128B).
    // The eight chunk offsets of this lane never change: keep them in registers so that the steady
    // state addresses shared memory with no per-access arithmetic.
    unsigned off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = (unsigned)cl * 128u + ((((unsigned)j) ^ (unsigned)(cl & 7)) << 4);

    for (int i = 0; i < n_tiles; ++i) {
        const int slot = i % St;
        const int t0 = i * tile_t;
        const int nt = a.n_samples - t0 < tile_t ? a.n_samples - t0 : tile_t;
        unsigned char* stage = my + (size_t)slot * stage_bytes;
        auto chunk_ptr = [&](int c) {                  // chunk c of the tile, this lane's channel
            return stage + ((unsigned)c >> 3) * kBoxBytes + (unsigned)cl * 128u +
                   ((((unsigned)c & 7u) ^ (unsigned)(cl & 7)) << 4);
        };

        mbar_wait(&bars[slot], (unsigned)((i / St) & 1));

        // `nxt` = the 4*CH inputs of the NEXT iteration, loaded one iteration ahead
        float4 nxt[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            nxt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (first) nxt[u] = *reinterpret_cast<const float4*>(chunk_ptr(u));
        }

        auto slow_iter = [&](int g) {                  // predicated: lanes may be outside the tile
            float4 cur[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) cur[u] = nxt[u];
            __syncwarp();
#pragma unroll
            for (int u = 0; u < CH; ++u) {                       // one LDS per lane, no divergence
                const unsigned char* src = x_in + u * 512;
                if (first && CH * (g + 1) + u < NB * 8) src = chunk_ptr(CH * (g + 1) + u);
                nxt[u] = *reinterpret_cast<const float4*>(src);
            }
            const int c = g - LAG * sec;               // this lane's position, in iterations
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int m = 4 * (CH * c + u) + q;
                    const bool act = ch_ok && m >= 0 && m < nt;
                    const float in = q == 0 ? cur[u].x : q == 1 ? cur[u].y : q == 2 ? cur[u].z : cur[u].w;
                    o[q] = f.eval(in);
                    f.push_if(act, in, o[q]);
                    if (last && act) reinterpret_cast<float*>(chunk_ptr(CH * c + u))[q] = o[q];
                }
                // (the slot of a channel's last lane is read by nobody: the lane above is a first lane)
                *reinterpret_cast<float4*>(x_out + u * 512) = make_float4(o[0], o[1], o[2], o[3]);
            }
        };
        // steady state, iteration IPB*box + j: chunks CH*j .. CH*j+CH-1 of `box` in (loaded one iteration
        // ahead: those of iteration j+1), the chunks of iteration j - DRAIN (one or two boxes back) out.
        // Every lane keeps the shared-memory address of each of its IPB loads and stores of a box in
        // registers: first / last lanes of a channel point into the tile (and move on by one box per box),
        // the others at their exchange slots (and stay) -- the steady state has no address selection left
        // and the LDS can issue right after the warp has synchronised.
        unsigned char* sptr[IPB][CH];
        unsigned char* dptr[IPB][CH];
        auto set_ptrs = [&](int b) {
            unsigned char* box = stage + (unsigned)b * kBoxBytes;
#pragma unroll
            for (int j = 0; j < IPB; ++j) {
                const int back = (DRAIN - j + IPB - 1) / IPB;                 // boxes back (>= 0)
                const int jo = (j - DRAIN + 8 * IPB) % IPB;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    // (the last iteration of a box prefetches from the next box; past the tile's last box
                    // that is some other shared memory of this CTA: read, never used)
                    sptr[j][u] = first ? (j < IPB - 1 ? box + off[(CH * (j + 1) + u) & 7] : box + kBoxBytes + off[u])
                                       : x_in + u * 512;
                    dptr[j][u] = last ? box - back * kBoxBytes + off[CH * jo + u] : x_out + u * 512;
                }
            }
        };
        const unsigned sinc = first ? (unsigned)kBoxBytes : 0u, dinc = last ? (unsigned)kBoxBytes : 0u;
        auto next_box = [&]() {
#pragma unroll
            for (int j = 0; j < IPB; ++j)
#pragma unroll
                for (int u = 0; u < CH; ++u) { sptr[j][u] += sinc; dptr[j][u] += dinc; }
        };
        auto fast_iter = [&](auto jc) {
            constexpr int j = decltype(jc)::value;
            float4 cur[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) cur[u] = nxt[u];
            __syncwarp();
#pragma unroll
            for (int u = 0; u < CH; ++u) nxt[u] = *reinterpret_cast<const float4*>(sptr[j][u]);
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float in = q == 0 ? cur[u].x : q == 1 ? cur[u].y : q == 2 ? cur[u].z : cur[u].w;
                    o[q] = f.eval(in);
                    f.push(in, o[q]);
                }
                *reinterpret_cast<float4*>(dptr[j][u]) = make_float4(o[0], o[1], o[2], o[3]);
            }
        };
        // iterations j0 .. IPB-1 of the current box (compile-time position inside the box)
        auto fast_span = [&](auto j0c) {
            constexpr int j0 = decltype(j0c)::value;
            if constexpr (j0 <= 0) fast_iter(std::integral_constant<int, 0>{});
            if constexpr (j0 <= 1 && IPB > 1) fast_iter(std::integral_constant<int, 1 % IPB>{});
            if constexpr (j0 <= 2 && IPB > 2) fast_iter(std::integral_constant<int, 2 % IPB>{});
            if constexpr (j0 <= 3 && IPB > 3) fast_iter(std::integral_constant<int, 3 % IPB>{});
            if constexpr (j0 <= 4 && IPB > 4) fast_iter(std::integral_constant<int, 4 % IPB>{});
            if constexpr (j0 <= 5 && IPB > 5) fast_iter(std::integral_constant<int, 5 % IPB>{});
            if constexpr (j0 <= 6 && IPB > 6) fast_iter(std::integral_constant<int, 6 % IPB>{});
            if constexpr (j0 <= 7 && IPB > 7) fast_iter(std::integral_constant<int, 7 % IPB>{});
        };

        const int nfull = nt >> 5;                     // boxes of this tile without a ragged tail
        const int total = (nt + 4 * CH - 1) / (4 * CH) + DRAIN;   // iterations until the last lane has drained
        constexpr int HB = (DRAIN + IPB - 1) / IPB;    // boxes the pipeline needs to fill
        if (nfull >= HB) {
            for (int g = 0; g < DRAIN; ++g) slow_iter(g);
            // rest of box HB-1: every lane is inside the tile from iteration DRAIN on
            set_ptrs(HB - 1);
            fast_span(std::integral_constant<int, DRAIN % IPB == 0 ? IPB : DRAIN % IPB>{});
#pragma unroll 1
            for (int b = HB; b < nfull; ++b) {
                next_box();
                fast_span(std::integral_constant<int, 0>{});
            }
            for (int g = nfull * IPB; g < total; ++g) slow_iter(g);
        } else {
            for (int g = 0; g < total; ++g) slow_iter(g);
        }

        fence_proxy_async();                           // generic-proxy writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
            int nbf; bool rag;
            tile_shape(i, nbf, rag);
            if (!rag) {
                tma_store_3d(&a.out_map[1], 0, c0, i * NB, stage);      // boxes past the end are clipped
            } else {
                for (int b = 0; b <= nbf; ++b)
                    tma_store_2d(&a.out_map[0], (i * NB + b) * kTileT, c0, stage + b * kBoxBytes);
            }
            tma_commit();
            const int nxt_tile = i + St - 1;
            if (nxt_tile < n_tiles) {
                tma_wait_read<1>();                    // the slot of tile i-1: its store has left smem
                issue_load(nxt_tile);
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
