// zignal-b200 :: biquad cascade with the SECTIONS of a channel group spread over the warps of a CTA (K1s).
//
// Same graph, same arithmetic, same state rows as zg_biquad.cuh (reference spelling test/benchmark.cpp:25-33,
// `fwd |= bwd` chained SECTIONS times).  Built for MANY channels, where the lane-per-channel kernel (K1) is limited by
// the length of the contiguous run each channel row contributes to an HBM request: K1 needs ~7 resident warps per SM to
// issue its arithmetic, every warp owns 32 channel rows, and 227 KB of shared memory over 7 x 32 rows x 2 stages is
// 512 bytes per row at a time (the skeleton streams 5.87 TB/s at 512 B, 6.23 TB/s at 1 KB) -- and
// since then the kernel of every channel count of the 4-section cascade (fewer groups per CTA for fewer channels, down
// to one group per SM with several boxes per hand-over, kHB below), also cut in time (FAST mode, SplitArgs::n_segs).
//
// Here a GROUP of WPG = SECTIONS / SPW warps shares one ring of tiles of 32 channel rows: warp `sec` of the group
// evaluates sections [sec * SPW, (sec + 1) * SPW) of those 32 channels (lane = channel, as in K1) IN PLACE, one
// 32-sample box behind warp sec - 1 -- the hand-over between sections is the tile itself.  Warps per SM and rows
// per SM are decoupled: three groups of four warps issue like twelve, but only 96 rows share the shared memory, so a
// tile is 8 boxes = 1 KB of every row in one TMA request.
//
//   * hand-over: one mbarrier per (section boundary, ring box); lane 0 of the producing warp arrives after a
//     __syncwarp (kAllArrive: every lane for itself), all lanes of the consuming warp wait.  Steady state: the wait
//     succeeds at once.
//   * TMA: one 3-D operation per tile and direction (box {32 samples, 32 channels, NB boxes}, SWIZZLE_128B, the same
//     shared-memory image as NB single boxes).  Lane 0 of the LAST warp of a group issues both: the store of the tile
//     it has just finished, and -- one box into the next tile, when that store has read its shared memory -- the load
//     of the tile S tiles ahead into the same stage.  The first warp waits for the load (mbarrier, tx bytes).
//   * persistent and balanced: grid = SMs, and every group works through a contiguous range of the tile sequence
//     (row 0 tiles 0.., row 1 tiles 0.., ...), all ranges the same number of tiles.  The tile stream runs straight
//     through row boundaries (the loads of the next row are in flight while this row drains), so pipeline fill and
//     drain happen once per launch.  65 536 channels are 2048 rows for 3 x 148 groups = 4.61 rows each: a range begins
//     and ends in the middle of a row, and the row's delay lines travel from the group that runs its head to the group
//     that runs its tail through HBM (`carry` + one flag per row and warp).  A group runs the head piece of its LAST
//     row FIRST and the tail piece of its first row LAST, so the group before it has long published the carry when it
//     is needed; CTAs number themselves in the order they start (a ticket), so the CTA waited for is always running.
//   * state: warp `sec` keeps the two-tick history of its input signal and of its SPW output signals (the history of
//     signal k is both the y-line of section k - 1 and the x-line of section k: two warps hold a copy, the producer
//     writes it back), read at the first box of a piece and written after its last -- after every warp of the group has
//     read its initial state (one more mbarrier per piece: the producer of a line may be a whole piece ahead of its
//     consumer).
//
// Every section is evaluated by BiquadDf1Cascade<SPW, ...>::tick -- the very code K1 runs -- so EXACT stays
// bit-identical; only which warp evaluates a section, and when, changes.
// Planar fp32 blocks with T % 32 == 0 only (the host falls back to K1 otherwise).
#pragma once
#include "zg_biquad.cuh"
#include "zg_biquad_lanes.cuh"

namespace zgk {

constexpr int kSplitAckRing = 8;           // pieces of work a group may have in flight (>= stages + 2)

struct SplitArgs {
    TensorMap in_map, out_map;              // 3-D whole-tile maps {32 samples, C channels, T/32 boxes}, box {32, 32, NB}
    float* state;                           // [n_state][ch_stride]
    const float* params;                    // [n_params][ch_stride]
    long long ch_stride;
    int channels;
    int n_samples;                          // a multiple of 32
    int boxes;                              // NB: boxes per tile
    int stages;                             // S >= 2 tiles in a group's ring
    unsigned long long* ctl;                // [0] tickets drawn: CTAs number themselves in the order they start;
                                            // [1] CTAs finished; [2] launches finished = the epoch before this launch.
                                            // The last CTA to finish zeroes [0], [1] and bumps [2]: nothing about a launch
                                            // lives on the host, so a captured launch can be replayed from a CUDA graph
    unsigned long long* flags;              // [channel group][warps per group]: the epoch of the launch whose head piece of
                                            //   this row is done
    float* carry;                           // [warps per group * state floats per warp][ch_stride]: a warp's delay lines
                                            //   between the two pieces of a row
    // Time segments (FAST mode, warm-up form; zg_runtime.cu choose_segments): with n_segs > 1 a "row" of the
    // tile sequence is (32 channels, segment g): boxes [g * seg_boxes, g * seg_boxes + seg_boxes + warm_boxes) of the block.
    // Segment 0 continues from the state rows; segment g > 0 starts from ZERO state and discards the outputs of its first
    // warm_boxes boxes (the host has checked that the graph forgets its state within that many ticks), so it answers for
    // the boxes from g * seg_boxes + warm_boxes on; the last segment takes the ragged end and leaves the block's final
    // state in state_out (another buffer than `state`: a first segment may still have to read that).
    int n_segs, seg_boxes, warm_boxes;      // seg_boxes and warm_boxes are multiples of `boxes`
    float* state_out;
    long long carry_stride;                 // floats per row of `carry`: ch_stride x n_segs
    int state_row[kMaxState];
    float uparams[kMaxUniform];
};

// mbarriers per group: S "tile landed" + (WPG - 1) x ring boxes "box handed over" + "every warp of the group has read
// the initial state of this piece" + S "the last warp is done with the tile that was in this stage"
__host__ __device__ constexpr int split_bar_count(int stages, int boxes, int wpg) {
    return stages + (wpg - 1) * stages * boxes + kSplitAckRing + stages;
}
__host__ __device__ constexpr int split_group_extra_bytes(int stages, int boxes, int wpg) {
    return 8 * split_bar_count(stages, boxes, wpg);
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void split_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// kAllArrive: every lane arrives on the hand-over barriers itself instead of lane 0 after a __syncwarp -- the form
// compute-sanitizer racecheck can follow (it tracks a thread's own arrivals only); 4.6 % slower in EXACT mode (32
// arrivals per box serialise), so it is built for the 4-section kernel only and chosen by ZG_TUNE_SPLIT_ARRIVE=1.
// kHB: boxes per hand-over.  1 for many channels (three groups per SM hide a hand-over behind each other's arithmetic).
// With few channels -- one group per SM, every warp alone on its scheduler -- the fixed cost of a hand-over (barrier
// round trip, first load of the box, drain of the stores) sits on the critical path of every box: kHB = 4 pays it once
// per four boxes (the warps of a group then run four boxes apart).
template <int SECTIONS, int SPW, bool kExact, bool kSym, bool kUniform, bool kAllArrive = false, int kHB = 1>
__device__ __forceinline__ void biquad_split_block(const SplitArgs& a) {
    static_assert(kHB == 1 || (SPW == 1 && !kSym), "runs of boxes are evaluated one plain section per warp");
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
    volatile unsigned long long* s_ctl = reinterpret_cast<volatile unsigned long long*>(
        tiles + (size_t)G * R * kTileBytes + (size_t)G * split_group_extra_bytes(S, NB, WPG));      // [0] ticket, [1] epoch
    unsigned long long* full = bars;                                       // [S] tile landed
    unsigned long long* ack_bar = bars + S + (size_t)(WPG - 1) * R;       // [kSplitAckRing]
    unsigned long long* empty = ack_bar + kSplitAckRing;                   // [S]
    constexpr unsigned lanes_arriving = kAllArrive ? 32u : 1u;
    // what this warp waits for before a box: the tile (first warp; one barrier per stage, re-checked per box: free
    // once the phase has completed) or the box from the warp before it; and what it signals after a box
    unsigned long long* wait_bar = first ? full : bars + S + (size_t)(sec - 1) * R;
    const int wait_per_box = first ? 0 : 1;
    unsigned long long* hand_out = bars + S + (size_t)(last ? 0 : sec) * R;   // [R] (unused by the last warp)

    if (sec == 0 && lane == 0) {
        // a tile lands once (one expect_tx arrival); a box / an initial state / a finished tile is handed over by lane 0 of a
        // warp after a __syncwarp, or by every lane for itself
        for (int i = 0; i < split_bar_count(S, NB, WPG); ++i) {
            unsigned long long* b = &bars[i];
            mbar_init(b, i < S ? 1u : b >= ack_bar && b < empty ? lanes_arriving * WPG : lanes_arriving);
        }
        fence_barrier_init();
        prefetch_tmap(&a.in_map);
        prefetch_tmap(&a.out_map);
    }
    // CTAs are numbered in the order they START: a CTA only ever waits for the one numbered before it (below), which is
    // therefore running or finished whatever else occupies the GPU and in whatever order the hardware starts CTAs
    if (threadIdx.x == 0) {
        s_ctl[0] = atomicAdd(&a.ctl[0], 1ull);
        s_ctl[1] = ld_acquire_gpu(&a.ctl[2]) + 1ull;    // stays put until every CTA of this launch has finished
    }
    __syncthreads();
    const unsigned long long epoch = s_ctl[1];

    // ---- the work of this group: a contiguous range of the tile sequence (row 0 tiles 0.., row 1 tiles 0.., ...) ----
    // Every group gets the same number of tiles (+-1), so a range begins and ends in the middle of a row.  The host
    // makes ranges at least one row long: a row is cut at most once.  Order inside the range: FIRST the head piece of
    // the row the range ends in (tiles [0, k_hi) of row_hi; leaves its delay lines in `carry` and raises the row's flag),
    // then the whole rows, LAST the tail piece of the row it begins in (tiles [k_lo, ..) of row_lo, continuing from the
    // carry of the group before -- which wrote it at the very start of the launch).
    const int n_cg = (a.channels + 31) >> 5;
    const int n_segs = a.n_segs > 1 ? a.n_segs : 1;
    const int block_boxes = a.n_samples / kTileT;
    // tiles per row; a row is a channel group, or (channel group, segment): row = cg * n_segs + g
    const int tpr = n_segs > 1 ? (a.seg_boxes + a.warm_boxes) / NB : (a.n_samples + tile_t - 1) / tile_t;
    const long long total = (long long)n_cg * n_segs * tpr;
    const int n_slots = (int)gridDim.x * G;
    const int slot = (int)s_ctl[0] * G + grp;
    const long long lo = total * slot / n_slots, hi = total * (slot + 1) / n_slots;
    const int row_lo = (int)(lo / tpr), k_lo = (int)(lo - (long long)row_lo * tpr);
    const int row_hi = (int)(hi / tpr), k_hi = (int)(hi - (long long)row_hi * tpr);
    const int has_head = k_hi > 0 && hi > lo ? 1 : 0;
    const int first_full = k_lo > 0 ? row_lo + 1 : row_lo;
    const int n_full = row_hi > first_full ? row_hi - first_full : 0;
    const int has_tail = k_lo > 0 && hi > lo ? 1 : 0;
    const int n_pieces = has_head + n_full + has_tail;
    auto piece = [&](int pi, int& row, int& kb, int& ke) {
        if (has_head) {
            if (pi == 0) { row = row_hi; kb = 0; ke = k_hi; return; }
            pi -= 1;
        }
        if (pi < n_full) { row = first_full + pi; kb = 0; ke = tpr; return; }
        row = row_lo; kb = k_lo; ke = tpr;
    };

    // ---- the load stream (lane 0 of the last warp): the tiles of the pieces in that order ----
    int ld_pi = 0, ld_row = 0, ld_k = 0, ld_ke = 0, ld_st = 0;
    bool ld_done = n_pieces == 0;
    if (!ld_done) piece(0, ld_row, ld_k, ld_ke);
    auto issue_load = [&]() {
        mbar_expect_tx(&full[ld_st], (unsigned)(NB * kTileBytes));
        const int ld_cg = ld_row / n_segs, ld_g = ld_row - ld_cg * n_segs;
        tma_load_3d(ring + (size_t)ld_st * NB * kTileBytes, &a.in_map, 0, ld_cg * 32, ld_g * a.seg_boxes + ld_k * NB, &full[ld_st]);
        if (++ld_st == S) ld_st = 0;
        if (++ld_k == ld_ke) {
            if (++ld_pi == n_pieces) ld_done = true;
            else piece(ld_pi, ld_row, ld_k, ld_ke);
        }
    };
    if (last && lane == 0) {
        for (int i = 0; i < S && !ld_done; ++i) issue_load();
    }

    // parameters of this warp's sections: shared ones once, per-channel ones at every piece
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

    // a warp signals: lane 0 for all lanes (their stores are ordered before its arrival by the __syncwarp), or -- for the
    // race checker, which follows a thread's own arrivals only -- every lane for itself
    auto warp_arrive = [&](unsigned long long* bar) {
        if constexpr (kAllArrive) split_arrive(bar);
        else {
            __syncwarp();
            if (lane == 0) split_arrive(bar);
        }
    };
    // 32 ticks of this warp's sections on one box, in place
    auto compute_box = [&](unsigned char* box) {
        {
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
    };
    // one box: wait until it is ours, evaluate it, hand it on
    auto do_box = [&](unsigned char* box, unsigned long long* wbar, unsigned par, unsigned long long* obar, bool compute) {
        mbar_wait(wbar, par);
        if (compute) compute_box(box);
        if (!last) warp_arrive(obar);
    };

    int st = 0, tile_no = 0;                           // stage / number of the next tile
    unsigned par = 0;
    bool any_tile = false;                             // a tile has been stored by this (last) warp: a stage to refill
    for (int pi = 0; pi < n_pieces; ++pi) {
        int row, kb, ke;
        piece(pi, row, kb, ke);
        const int cg = row / n_segs, seg = row - cg * n_segs;
        const int c0 = cg * 32;
        const int ch = c0 + lane;
        const bool ch_ok = ch < a.channels;
        const bool from_carry = kb > 0, to_carry = ke < tpr;
        const int box_first = seg * a.seg_boxes;                            // of the row, in boxes of the block
        const int box_end = n_segs > 1 && box_first + a.seg_boxes + a.warm_boxes < block_boxes
                                ? box_first + a.seg_boxes + a.warm_boxes : block_boxes;
        const int k_keep = seg > 0 ? a.warm_boxes / NB : 0;                 // tiles before this one produce state only
        unsigned long long* my_flag = a.flags + (size_t)row * WPG + sec;
        unsigned long long* ack = &ack_bar[pi & (kSplitAckRing - 1)];
        const long long carry_col = (long long)row * 32 + lane;             // (n_segs == 1: the channel)

        // its delay lines (and coefficients): from the state rows (a later segment: zero), or -- the tail piece of a row --
        // from what the same warp of the group before left in `carry` (ld.cg: written by another SM during this launch)
        if (from_carry) {
            // (the group before this one published the carry at the very start of the launch, so this wait is over before
            //  it begins; should it ever not end -- a planner bug: nobody runs that head piece -- the launch fails with an
            //  error after ~10 s of polling instead of hanging the GPU)
            for (unsigned polls = 0; ld_acquire_gpu(my_flag) != epoch;)
                if (++polls > (1u << 24)) __trap();
#pragma unroll
            for (int j = 0; j < NS; ++j) s[j] = ch_ok ? __ldcg(&a.carry[(long long)(sec * NS + j) * a.carry_stride + carry_col]) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j)
                s[j] = ch_ok && seg == 0 ? a.state[(long long)a.state_row[2 * SPW * sec + j] * a.ch_stride + ch] : 0.f;
        }
        // the state rows of a signal are read by the warp that consumes it and written by the warp that produces it,
        // which may be a whole piece ahead: nobody writes before everybody has read
        warp_arrive(ack);
        if (!kUniform) {
#pragma unroll
            for (int j = 0; j < NP; ++j) prm[j] = ch_ok ? a.params[(long long)(sec * NP + j) * a.ch_stride + ch] : 0.f;
        }
        if constexpr (NE > 0) Tick::init(s, prm, ex);

        for (int k = kb; k < ke; ++k) {
            const int box0 = box_first + k * NB;       // first box of the tile, in boxes of the block
            const int left = box_end - box0;
            const int nb = left < NB ? (left > 0 ? left : 0) : NB;      // boxes of this tile that hold samples
            unsigned char* stage = ring + (size_t)st * NB * kTileBytes;
            unsigned long long* wb = wait_bar + (first ? st : st * NB);
            unsigned long long* ob = hand_out + st * NB;

            // (the load into this stage was issued after the store of the tile before it had read the stage, so this wait
            //  never blocks: it states, for the memory model and the race checker, that the last warp's stores to the
            //  stage happened before the first warp's loads from it S tiles later)
            if (first && tile_no >= S) mbar_wait(&empty[st], par ^ 1u);
            if constexpr (kHB > 1) {
                // boxes in runs of kHB (NB is a multiple of kHB, and at least two runs): one wait and one arrival per run
#pragma unroll 1
                for (int b0 = 0; b0 < NB; b0 += kHB) {
                    mbar_wait(wb + b0 * wait_per_box, par);
                    if (b0 == kHB && last && lane == 0 && any_tile && !ld_done) {
                        tma_wait_read<0>();
                        issue_load();
                    }
                    // the ticks of the run through Df1Lane::step4 (zg_biquad_lanes.cuh): the recurrence
                    // y = (v + a1*y1) + a2*y2 -- three dependent instructions per sample, all a lone warp has to wait
                    // for -- with the feed-forward half of the NEXT four samples computed in its shadow.  Same
                    // operations in the same association as the tick: bit-identical.
                    const int cnt = nb - b0 < kHB ? nb - b0 : kHB;
                    if (cnt > 0) {
                        Df1Lane<kExact> f;
                        f.b0 = prm[0]; f.b1 = prm[1]; f.b2 = prm[2]; f.a1 = prm[3]; f.a2 = prm[4];
                        f.x2 = s[0]; f.x1 = s[1]; f.y2 = s[2]; f.y1 = s[3];
                        unsigned char* bx = stage + (size_t)b0 * kTileBytes;
                        float4 c = *reinterpret_cast<const float4*>(bx + offs[0]);
                        float v[4], o[4];
                        f.prime(c, v);
#pragma unroll
                        for (int h = 0; h < kHB; ++h) {
                            if (h < cnt) {
                                unsigned char* box = bx + (size_t)h * kTileBytes;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float4 n = c;
                                    if (j < 7) n = *reinterpret_cast<const float4*>(box + offs[j + 1]);
                                    else if (h + 1 < cnt) n = *reinterpret_cast<const float4*>(box + kTileBytes + offs[0]);
                                    f.step4(c, n, v, o);
                                    *reinterpret_cast<float4*>(box + offs[j]) = make_float4(o[0], o[1], o[2], o[3]);
                                    c = n;
                                }
                            }
                        }
                        s[0] = f.x2; s[1] = f.x1; s[2] = f.y2; s[3] = f.y1;
                    }
                    if (!last) warp_arrive(ob + b0);
                }
            } else {
            do_box(stage, wb, par, ob, nb > 0);         // (a segment's last tiles may lie past the end of the block)
            if (last && lane == 0 && any_tile && !ld_done) {
                tma_wait_read<0>();                    // the store of the previous tile has read its stage:
                issue_load();                          // the tile S - 1 tiles ahead goes there
            }
#pragma unroll 1
            for (int b = 1; b < nb; ++b) do_box(stage + (size_t)b * kTileBytes, wb + b * wait_per_box, par, ob + b, true);
#pragma unroll 1
            for (int b = nb > 1 ? nb : 1; b < NB; ++b) do_box(stage, wb + b * wait_per_box, par, ob + b, false);    // past the end of the row
            }

            if (last) {
                fence_proxy_async();                   // generic-proxy writes -> visible to TMA
                __syncwarp();
                if (lane == 0) {
                    if (k >= k_keep) tma_store_3d(&a.out_map, 0, c0, box0, stage);      // (warm-up tiles: state, not samples)
                    tma_commit();
                }
                any_tile = true;
                warp_arrive(&empty[st]);
            }
            ++tile_no;
            if (++st == S) { st = 0; par ^= 1u; }
        }

        // ---- the piece is finished: its delay lines to the group that continues the row, or -- at the end of the row --
        //      back to the state rows (every line by the warp that produces it) ----
        mbar_wait(ack, (unsigned)((pi / kSplitAckRing) & 1));
        if (to_carry) {
            if (ch_ok) {
#pragma unroll
                for (int j = 0; j < NS; ++j) a.carry[(long long)(sec * NS + j) * a.carry_stride + carry_col] = s[j];
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                st_release_gpu(my_flag, epoch);
            }
        } else if (ch_ok && seg == n_segs - 1) {       // the end of the block for these channels
#pragma unroll
            for (int j = 0; j < NS; ++j)
                if (j >= 2 || first) a.state_out[(long long)a.state_row[2 * SPW * sec + j] * a.ch_stride + ch] = s[j];
        }
    }
    if (last && lane == 0) tma_wait_all<0>();           // shared memory must outlive the last stores

    // the last CTA to finish leaves the counters ready for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&a.ctl[1], 1ull) + 1ull == gridDim.x) {
            a.ctl[0] = 0ull;
            a.ctl[1] = 0ull;
            __threadfence();
            st_release_gpu(&a.ctl[2], epoch);
        }
    }
}

}  // namespace zgk
