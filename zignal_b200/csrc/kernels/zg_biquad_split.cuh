// zignal-b200 :: biquad cascade with the SECTIONS of a channel group spread over the warps of a CTA (K1s).
//
// Same graph, same arithmetic, same state rows as zg_biquad.cuh (reference spelling test/benchmark.cpp:25-33,
// `fwd |= bwd` chained SECTIONS times); for MANY channels, where the lane-per-channel kernel (K1) is limited by the
// length of the contiguous run each channel row contributes to an HBM request: K1 needs ~7 resident warps per SM to
// issue its arithmetic, every warp owns 32 channel rows, and 227 KB of shared memory over 7 x 32 rows x 2 stages is
// 512 bytes per row at a time (the skeleton streams 5.87 TB/s at 512 B, 6.23 TB/s at 1 KB).
//
// Here a GROUP of WPG = SECTIONS / SPW warps shares one ring of tiles of 32 channel rows: warp `sec` of the group
// evaluates sections [sec * SPW, (sec + 1) * SPW) of those 32 channels (lane = channel, as in K1) IN PLACE, one
// 32-sample box behind warp sec - 1 -- the hand-over between sections is the tile itself.  Warps per SM and rows
// per SM are decoupled: two groups of four warps issue like eight, but only 64 rows share the shared memory, so a
// tile is 12-14 boxes = 1.5-1.75 KB of every row in one TMA request.
//
//   * hand-over: one mbarrier per (section boundary, ring box); lane 0 of the producing warp arrives after a
//     __syncwarp, all lanes of the consuming warp wait.  Steady state: the wait succeeds at once.
//   * TMA: one 3-D operation per tile and direction (box {32 samples, 32 channels, NB boxes}, SWIZZLE_128B, the same
//     shared-memory image as NB single boxes).  Lane 0 of the LAST warp of a group issues both: the store of the tile
//     it has just finished, and -- one box into the next tile, when that store has read its shared memory -- the load
//     of the tile S tiles ahead into the same stage.  The first warp waits for the load (mbarrier, tx bytes).
//   * persistent: grid = SMs; a group walks the channel groups slot, slot + n_slots, ... and the tile stream runs
//     straight through row boundaries (the loads of the next row's first tiles are in flight while the last tiles of
//     this row drain), so pipeline fill and drain happen once per launch, not once per row.
//   * state: warp `sec` keeps the two-tick history of its input signal and of its SPW output signals (the history of
//     signal k is both the y-line of section k - 1 and the x-line of section k: two warps hold a copy, the producer
//     writes it back), read at the first box of a row and written after its last.
//
// Every section is evaluated by BiquadDf1Cascade<SPW, ...>::tick -- the very code K1 runs -- so EXACT stays
// bit-identical; only which warp evaluates a section, and when, changes.
// Planar fp32 blocks with T % 32 == 0 only (the host falls back to K1 otherwise).
#pragma once
#include "zg_biquad.cuh"
#include "zg_biquad_lanes.cuh"

namespace zgk {

// bytes of mbarriers per group: S "tile landed" + (WPG - 1) x ring boxes "box handed over"
__host__ __device__ constexpr int split_bar_count(int stages, int boxes, int wpg) {
    return stages + (wpg - 1) * stages * boxes;
}

template <int SECTIONS, int SPW, bool kExact, bool kSym, bool kUniform>
__device__ __forceinline__ void biquad_split_block(const StreamArgs& a) {
    static_assert(SECTIONS % SPW == 0, "sections per warp must divide the cascade");
    constexpr int WPG = SECTIONS / SPW;                        // warps per group
    typedef BiquadDf1Cascade<SPW, kExact, kSym> Tick;
    constexpr int NS = Tick::N_STATE, NP = Tick::N_PARAM, NE = extra_count<Tick>::value;

    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    // through a shuffle so that the compiler knows it is warp-uniform: the ring / barrier addresses derived from it
    // then live in uniform registers and the role branches are uniform branches
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp = warp / WPG, sec = warp - grp * WPG;
    const int G = (int)(blockDim.x >> 5) / WPG;
    const int S = a.stages, NB = a.boxes, R = S * NB;
    const int tile_t = NB * kTileT;
    const bool first = sec == 0, last = sec == WPG - 1;

    unsigned char* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char* ring = tiles + (size_t)grp * R * kTileBytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + (size_t)G * R * kTileBytes) +
                               (size_t)grp * split_bar_count(S, NB, WPG);
    unsigned long long* full = bars;                                       // [S] tile landed
    // what this warp waits for before a box: the tile (first warp; one barrier per stage, re-checked per box: free
    // once the phase has completed) or the box from the warp before it; and what it signals after a box
    unsigned long long* wait_bar = first ? full : bars + S + (size_t)(sec - 1) * R;
    const int wait_per_box = first ? 0 : 1;
    unsigned long long* hand_out = bars + S + (size_t)(last ? 0 : sec) * R;   // [R] (unused by the last warp)

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < G * split_bar_count(S, NB, WPG); ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
        prefetch_tmap(&a.in_map[1]);
        prefetch_tmap(&a.out_map[1]);
    }
    __syncthreads();

    // which channel groups this group walks, and the tile stream through them
    const int n_cg = (a.channels + 31) >> 5;
    const int n_slots = (int)gridDim.x * G;
    const int slot = (int)blockIdx.x * G + grp;
    const int tiles_per_row = (a.n_samples + tile_t - 1) / tile_t;
    const int my_rows = slot < n_cg ? (n_cg - slot + n_slots - 1) / n_slots : 0;
    const int n_tiles = my_rows * tiles_per_row;

    // the load stream (lane 0 of the last warp): tiles in order, tile ld_i = tile ld_k of my row ld_r into stage ld_st
    int ld_i = 0, ld_r = 0, ld_k = 0, ld_st = 0;
    auto issue_load = [&]() {
        mbar_expect_tx(&full[ld_st], (unsigned)(NB * kTileBytes));
        tma_load_3d(ring + (size_t)ld_st * NB * kTileBytes, &a.in_map[1], 0, (slot + ld_r * n_slots) * 32, ld_k * NB, &full[ld_st]);
        ++ld_i;
        if (++ld_k == tiles_per_row) { ld_k = 0; ++ld_r; }
        if (++ld_st == S) ld_st = 0;
    };
    if (last && lane == 0) {
        const int pre = n_tiles < S ? n_tiles : S;
        for (int i = 0; i < pre; ++i) issue_load();
    }

    // parameters of this warp's sections: shared ones once, per-channel ones at every row
    Arr<NP> prm;
    if (kUniform) {
#pragma unroll
        for (int j = 0; j < NP; ++j) prm[j] = a.uparams[sec * NP + j];
    }
    Arr<NS> s;
    Arr<NE> ex;
    // row `lane` of a box, 16-byte chunk j at chunk (j ^ (lane & 7)) (SWIZZLE_128B): eight per-lane offsets for the
    // whole launch, added to the (uniform) address of the box
    unsigned offs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) offs[j] = (unsigned)lane * 128u + ((((unsigned)j) ^ (unsigned)(lane & 7)) << 4);

    // one box: wait until it is ours, 32 ticks of this warp's sections in place, hand it on
    auto do_box = [&](unsigned char* box, unsigned long long* wbar, unsigned par, unsigned long long* obar, bool compute) {
        mbar_wait(wbar, par);
        if (compute) {
            uint4 xn = *reinterpret_cast<const uint4*>(box + offs[0]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float xv[4], yv[4];
                Io<4>::unpack(xn, xv);
                if (j < 7) xn = *reinterpret_cast<const uint4*>(box + offs[j + 1]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    Arr<1> x, y;
                    x[0] = xv[q];
                    if constexpr (NE > 0) Tick::tick(x, y, s, prm, ex);
                    else Tick::tick(x, y, s, prm);
                    yv[q] = y[0];
                }
                *reinterpret_cast<uint4*>(box + offs[j]) = Io<4>::pack(yv);
            }
        }
        if (!last) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(obar)) : "memory");
        }
    };

    int r = 0, k = 0, st = 0;                          // tile i = tile k of my row r, in stage st
    unsigned par = 0;
    for (int i = 0; i < n_tiles; ++i) {
        const int t0 = k * tile_t;
        const int left = (a.n_samples - t0) / kTileT;
        const int nb = left < NB ? left : NB;          // boxes of this tile that hold samples
        const int c0 = (slot + r * n_slots) * 32;
        const int ch = c0 + lane;
        const bool ch_ok = ch < a.channels;
        unsigned char* stage = ring + (size_t)st * NB * kTileBytes;
        unsigned long long* wb = wait_bar + (first ? st : st * NB);
        unsigned long long* ob = hand_out + st * NB;

        if (k == 0) {                                  // a new row: its delay lines (and coefficients)
#pragma unroll
            for (int j = 0; j < NS; ++j)
                s[j] = ch_ok ? a.state[(long long)a.state_row[2 * SPW * sec + j] * a.ch_stride + ch] : 0.f;
            if (!kUniform) {
#pragma unroll
                for (int j = 0; j < NP; ++j) prm[j] = ch_ok ? a.params[(long long)(sec * NP + j) * a.ch_stride + ch] : 0.f;
            }
            if constexpr (NE > 0) Tick::init(s, prm, ex);
        }

        do_box(stage, wb, par, ob, true);
        if (last && lane == 0 && i >= 1 && ld_i < n_tiles) {
            tma_wait_read<0>();                        // the store of tile i-1 has read its stage:
            issue_load();                              // tile i-1+S goes there
        }
#pragma unroll 1
        for (int b = 1; b < nb; ++b) do_box(stage + (size_t)b * kTileBytes, wb + b * wait_per_box, par, ob + b, true);
#pragma unroll 1
        for (int b = nb; b < NB; ++b) do_box(stage, wb + b * wait_per_box, par, ob + b, false);    // past the end of the row

        if (last) {
            fence_proxy_async();                       // generic-proxy writes -> visible to TMA
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&a.out_map[1], 0, c0, k * NB, stage);
                tma_commit();
            }
        }

        if (++k == tiles_per_row) {                    // the row is finished: its delay lines back to HBM
            if (ch_ok) {
#pragma unroll
                for (int j = 0; j < NS; ++j)
                    if (j >= 2 || first) a.state[(long long)a.state_row[2 * SPW * sec + j] * a.ch_stride + ch] = s[j];
            }
            k = 0;
            ++r;
        }
        if (++st == S) { st = 0; par ^= 1u; }
    }
    if (last && lane == 0) tma_wait_all<0>();           // shared memory must outlive the last stores
}

}  // namespace zgk
