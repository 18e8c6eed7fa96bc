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
    // feed-forward half of a tick: (b0*x + b1*xa) + b2*xb, xa / xb = the inputs one / two ticks before x
    __device__ __forceinline__ float ff(float x, float xa, float xb) const {
        if (kExact) return __fadd_rn(__fadd_rn(__fmul_rn(b0, x), __fmul_rn(b1, xa)), __fmul_rn(b2, xb));
        return fmaf(b2, xb, fmaf(b1, xa, b0 * x));
    }
    // v[] of the four samples in c, from the plain state (start of a run of steady-state iterations)
    __device__ __forceinline__ void prime(const float4& c, float (&v)[4]) const {
        v[0] = ff(c.x, x1, x2); v[1] = ff(c.y, c.x, x1); v[2] = ff(c.z, c.y, c.x); v[3] = ff(c.w, c.z, c.y);
    }
    // Four ticks, software-pipelined by one iteration: the feed-forward halves v[] of the samples in c were
    // computed during the previous iteration; this one runs the recurrence  y = (v + a1*y1) + a2*y2  -- three
    // dependent 4-cycle instructions per sample, the critical path of the whole kernel -- and computes, in
    // its shadow, v[] of the NEXT four samples n (their history is c.z, c.w).  Same operations in the same
    // association as eval(): bit-identical.  The statements are interleaved the way the instruction stream
    // should be (one recurrence instruction, two independent ones): with a single warp per scheduler
    // nothing else hides the 4-cycle latency.
    __device__ __forceinline__ void step4(const float4& c, const float4& n, float (&v)[4], float (&o)[4]) {
        if (kExact) {
            float t, s, w0, w1, w2, w3, m;
            const float p0 = __fmul_rn(a2, y2), p1 = __fmul_rn(a2, y1);
            t = __fmul_rn(a1, y1);      w0 = __fmul_rn(b0, n.x);  m = __fmul_rn(b1, c.w);
            s = __fadd_rn(v[0], t);     w0 = __fadd_rn(w0, m);    m = __fmul_rn(b2, c.z);
            o[0] = __fadd_rn(s, p0);    w0 = __fadd_rn(w0, m);    w1 = __fmul_rn(b0, n.y);
            t = __fmul_rn(a1, o[0]);    m = __fmul_rn(b1, n.x);   const float p2 = __fmul_rn(a2, o[0]);
            s = __fadd_rn(v[1], t);     w1 = __fadd_rn(w1, m);    m = __fmul_rn(b2, c.w);
            o[1] = __fadd_rn(s, p1);    w1 = __fadd_rn(w1, m);    w2 = __fmul_rn(b0, n.z);
            t = __fmul_rn(a1, o[1]);    m = __fmul_rn(b1, n.y);   const float p3 = __fmul_rn(a2, o[1]);
            s = __fadd_rn(v[2], t);     w2 = __fadd_rn(w2, m);    m = __fmul_rn(b2, n.x);
            o[2] = __fadd_rn(s, p2);    w2 = __fadd_rn(w2, m);    w3 = __fmul_rn(b0, n.w);
            t = __fmul_rn(a1, o[2]);    m = __fmul_rn(b1, n.z);
            s = __fadd_rn(v[3], t);     w3 = __fadd_rn(w3, m);    m = __fmul_rn(b2, n.y);
            o[3] = __fadd_rn(s, p3);    w3 = __fadd_rn(w3, m);
            v[0] = w0; v[1] = w1; v[2] = w2; v[3] = w3;
        } else {
            float s, w0, w1, w2, w3;
            s = fmaf(a1, y1, v[0]);     w0 = b0 * n.x;            w0 = fmaf(b1, c.w, w0);
            o[0] = fmaf(a2, y2, s);     w0 = fmaf(b2, c.z, w0);   w1 = b0 * n.y;
            s = fmaf(a1, o[0], v[1]);   w1 = fmaf(b1, n.x, w1);   w1 = fmaf(b2, c.w, w1);
            o[1] = fmaf(a2, y1, s);     w2 = b0 * n.z;            w2 = fmaf(b1, n.y, w2);
            s = fmaf(a1, o[1], v[2]);   w2 = fmaf(b2, n.x, w2);   w3 = b0 * n.w;
            o[2] = fmaf(a2, o[0], s);   w3 = fmaf(b1, n.z, w3);   w3 = fmaf(b2, n.y, w3);
            s = fmaf(a1, o[2], v[3]);
            o[3] = fmaf(a2, o[1], s);
            v[0] = w0; v[1] = w1; v[2] = w2; v[3] = w3;
        }
        x2 = c.z; x1 = c.w; y2 = o[2]; y1 = o[3];
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

constexpr int kLanesXbufBytes = 512;         // exchange buffer per warp: 32 lanes x 16 bytes

// in_map[0] / out_map[0]: 2-D {T, C}, box {32, 32/S}        (ragged last tile, box by box)
// in_map[1] / out_map[1]: 3-D {32, C, T/32} over the full 32-sample boxes, box {32, 32/S, NB}
template <int S, bool kExact, bool kUniform>
__device__ __forceinline__ void biquad_lanes_block(const StreamArgs& a) {
    static_assert(S == 2 || S == 4, "lanes per channel: 2 or 4 (the box must span whole swizzle atoms)");
    constexpr int CPW = 32 / S;              // channels per warp
    constexpr int kBoxBytes = CPW * 128;
    constexpr int PD = 3;                    // iterations between loading a chunk and evaluating it (its
                                             // feed-forward half is computed one iteration early: step4)
    constexpr int LAG = PD + 1;              // iterations between neighbouring sections
    constexpr int DRAIN = LAG * (S - 1);     // iterations until the last section has caught up
    constexpr int HB = (DRAIN + 7) / 8;      // boxes of iterations during which the pipeline fills

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
    // the St stages of a warp are contiguous: a ring of RB = St*NB boxes, global box Bx at (Bx mod RB)
    unsigned char* my = tiles + (size_t)warp * St * stage_bytes;
    const int RB = St * NB;
    unsigned char* xbuf = tiles + (size_t)warps_per_cta * St * stage_bytes + (size_t)warp * kLanesXbufBytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
                                   tiles + (size_t)warps_per_cta * (St * stage_bytes + kLanesXbufBytes)) + warp * St;
    // Exchange slots.  An LDS.128 / STS.128 is served a quarter-warp (8 lanes x 16 bytes) at a time and is
    // conflict free iff those 8 accesses fall into 8 different 16-byte bank groups.  In every quarter-warp the
    // first (last) lanes of its 8/S channels read (write) the tile: rows cl, chunk jj ^ (cl & 7) -- an aligned
    // block of 8/S bank groups that moves with jj -- while the other lanes read (write) exchange slots.  So
    // the slots of a quarter-warp live in one 128-byte line and are XOR-ed away from the block the tile
    // accesses of the same instruction use: `jj` = the chunk index the first lanes load in the iteration
    // that READS the slot (the last lanes' store of the writing iteration uses the same block: DRAIN - PD - 1
    // is a multiple of 8).  With plain lane-indexed slots every second wavefront was a bank conflict and the
    // four warps of an SM kept the shared-memory pipe busy 64 of ~76 cycles per iteration.
    constexpr int GB = 8 / S;                          // bank groups per block = channels per quarter-warp
    constexpr int LG = S == 4 ? 1 : 2;                 // log2(GB)
    static_assert((DRAIN - PD - 1) % 8 == 0, "slot swizzle assumes stores and loads use the same tile block");
    auto xslot = [&](int wl, int jj) {                 // slot written by lane wl (not a last lane)
        const int q = wl >> 3, ca = (wl / S) & (GB - 1), sw = wl & (S - 1);
        const int g0 = GB * (1 + sw) + ca;
        const int pb = ((jj & 7) >> LG) ^ (q & (S - 1));
        return xbuf + 128 * q + 16 * (g0 ^ (GB * pb));
    };

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
    *reinterpret_cast<float4*>(xbuf + lane * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const int T = a.n_samples;
    const int NC = (T + 3) >> 2;                       // 16-byte chunks (of four samples) of the block
    const int total = NC + DRAIN;                      // iterations: lane `sec` evaluates chunk G - LAG*sec in iteration G
    const int iter_boxes = (total + 7) >> 3;
    const int CPT = NB * 8;                            // chunks per tile
    const int n_tiles = (T + tile_t - 1) / tile_t;
    const int full_boxes = T >> 5;                     // boxes of the block without a ragged tail
    const bool ragged = (T & 31) != 0;
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
    auto issue_store = [&](int i) {                    // lane 0 only
        unsigned char* src = my + (size_t)(i % St) * stage_bytes;
        int nbf; bool rag;
        tile_shape(i, nbf, rag);
        if (!rag) {
            tma_store_3d(&a.out_map[1], 0, c0, i * NB, src);           // boxes past the end are clipped
        } else {
            for (int b = 0; b <= nbf; ++b)
                tma_store_2d(&a.out_map[0], (i * NB + b) * kTileT, c0, src + b * kBoxBytes);
        }
        tma_commit();
    };
    if (lane == 0) {
        const int pre = n_tiles < St ? n_tiles : St;
        for (int i = 0; i < pre; ++i) issue_load(i);
    }

    // row `cl` of every box; 16-byte chunk j of the row lives at chunk j ^ (cl & 7)   (SWIZZLE_128B)
    unsigned off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = (unsigned)cl * 128u + ((((unsigned)j) ^ (unsigned)(cl & 7)) << 4);
    auto ring_box = [&](int bx) { return my + (unsigned)(bx % RB) * kBoxBytes; };          // bx >= 0
    auto chunk_ptr = [&](int c) {                      // chunk c >= 0 of the block, this lane's channel
        return ring_box(c >> 3) + (unsigned)cl * 128u + ((((unsigned)c & 7u) ^ (unsigned)(cl & 7)) << 4);
    };

    mbar_wait(&bars[0], 0);                            // tile 0
    // r0 / r1 / r2: the inputs of this lane's next three iterations (loaded PD iterations ahead)
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    if (first) {
        r0 = *reinterpret_cast<const float4*>(chunk_ptr(0));
        r1 = *reinterpret_cast<const float4*>(chunk_ptr(1));
        r2 = *reinterpret_cast<const float4*>(chunk_ptr(2));
    }

    // fill, drain and ragged ends: any lane may be outside the block (predicated, branch-free, dynamic addresses)
    auto slow_iter = [&](int g) {
        const float4 cur = r0;
        r0 = r1;
        r1 = r2;
        __syncwarp();
        const unsigned char* src = first ? xbuf : xslot(lane - 1, g + PD);
        if (first && g + PD < n_tiles * CPT) src = chunk_ptr(g + PD);
        r2 = *reinterpret_cast<const float4*>(src);
        const int c = g - LAG * sec;                   // this lane's chunk
        float* dst = reinterpret_cast<float*>(chunk_ptr(c > 0 ? c : 0));
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool act = ch_ok && c >= 0 && 4 * c + q < T;
            const float in = q == 0 ? cur.x : q == 1 ? cur.y : q == 2 ? cur.z : cur.w;
            o[q] = f.eval(in);
            f.push_if(act, in, o[q]);
            if (last && act) dst[q] = o[q];
        }
        __syncwarp();
        if (!last) *reinterpret_cast<float4*>(xslot(lane, g + 1 + PD)) = make_float4(o[0], o[1], o[2], o[3]);
    };

    // Steady state.  Every lane keeps the shared-memory address of each of its 8 loads and 8 stores of a
    // box of iterations in registers: first / last lanes of a channel point into the ring (and move on by
    // one box per box), the others at their exchange slots (and stay), so an iteration is one LDS.128, four
    // ticks and one STS.128 with no address arithmetic.  In iteration 8B + j a first lane loads chunk
    // 8B + j + PD, a last lane stores chunk 8B + j - DRAIN.
    unsigned char* sptr[8];
    unsigned char* dptr[8];
    int bmod = 0;                                      // B mod RB, kept incrementally (no division in the loop)
    const unsigned sinc = first ? (unsigned)kBoxBytes : 0u, dinc = last ? (unsigned)kBoxBytes : 0u;
    auto set_ptrs_prev = [&]() {                       // the table of box B (ring position bmod), minus one step
        unsigned char* rb[4];                          // ring boxes B-2 .. B+1
#pragma unroll
        for (int d = -2; d <= 1; ++d) {
            int r = bmod + d;
            r = r < 0 ? r + RB : (r >= RB ? r - RB : r);
            rb[d + 2] = my + (unsigned)r * kBoxBytes;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int back = (DRAIN - j + 7) / 8;                      // boxes back (0 .. 2)
            const int jo = (j - DRAIN + 64) % 8;
            sptr[j] = first ? rb[2 + (j + PD) / 8] + off[(j + PD) & 7] - kBoxBytes : xslot(lane - 1, j + PD);
            dptr[j] = last ? rb[2 - back] + off[jo] - kBoxBytes : xslot(lane, j + 1 + PD);
        }
    };
    float v[4];                                        // feed-forward halves of the samples in r0 (see step4)
    auto fast_iter = [&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const float4 cur = r0;
        r0 = r1;
        r1 = r2;
        __syncwarp();
        r2 = *reinterpret_cast<const float4*>(sptr[j]);
        float o[4];
        f.step4(cur, r0, v, o);
        __syncwarp();                                  // the hand-over loads of all lanes come before any lane's store
                                                       // (one slot per lane: racecheck-clean; measured cost 0.4 %)
        *reinterpret_cast<float4*>(dptr[j]) = make_float4(o[0], o[1], o[2], o[3]);
    };

    // The pipeline runs through the whole block: tiles only matter to TMA.  Two events per tile, each due
    // before a known box of iterations; between events the hot loop is a run of boxes with nothing but a
    // counter:
    //   store  tile i is complete once the last lane has written its last chunk, DRAIN iterations after its
    //          last input chunk was read (two boxes into tile i+1).  At the same point the slot of tile i-1,
    //          stored a whole tile ago, is refilled with tile i-1+St (needs St >= 3: the host sees to it);
    //   wait   tile i must have landed before box i*NB - 1 (the prefetch of a first lane reaches into it).
    constexpr int kNever = 0x7fffffff;
    auto store_box = [&](int i) {                      // first box by which tile i is complete
        const int end_chunk = (i + 1) * CPT < NC ? (i + 1) * CPT : NC;
        return ((end_chunk + DRAIN - 1) >> 3) + 1;
    };
    int s_tile = 0, s_box = store_box(0);
    int r_tile = -1;                                   // tile to load into the slot stored at the previous event
    int w_tile = 2, w_box = n_tiles > 2 ? 2 * NB - 1 : kNever;
    if (n_tiles > 1) mbar_wait(&bars[1 % St], 0);      // tile 1 (the fill iterations may reach into it)
    auto do_store = [&]() {
        fence_proxy_async();                           // generic-proxy writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
            if (r_tile >= 0) {
                tma_wait_read<0>();
                issue_load(r_tile);
            }
            issue_store(s_tile);
        }
        r_tile = s_tile + St < n_tiles ? s_tile + St : -1;
        ++s_tile;
        s_box = s_tile < n_tiles ? store_box(s_tile) : kNever;
    };
    bool ptrs_valid = false;
    int B = 0;
#pragma unroll 1
    while (B < iter_boxes) {
        while (B >= s_box) do_store();
        while (B >= w_box) {
            mbar_wait(&bars[w_tile % St], (unsigned)((w_tile / St) & 1));
            ++w_tile;
            w_box = w_tile < n_tiles ? w_tile * NB - 1 : kNever;
        }
        if (B >= HB && B < full_boxes) {               // every lane is inside the block for all 8 iterations
            int stop = s_box < w_box ? s_box : w_box;
            stop = stop < full_boxes ? stop : full_boxes;
            int n = stop - B;
            // a pointer table survives a step of one box unless one of the boxes it refers to (B-2 .. B+1)
            // wraps around the ring: ring positions RB-1, 0, 1, 2 rebuild it
            if (!ptrs_valid || bmod <= 2 || bmod == RB - 1) {
                set_ptrs_prev();
                ptrs_valid = true;
                n = 1;
            } else {
                const int room = RB - 1 - bmod;
                n = n < room ? n : room;
            }
            B += n;
            bmod += n;
            bmod = bmod >= RB ? bmod - RB : bmod;
            f.prime(r0, v);
#pragma unroll 1
            for (; n > 0; --n) {                       // the hot loop
#pragma unroll
                for (int j = 0; j < 8; ++j) { sptr[j] += sinc; dptr[j] += dinc; }
                fast_iter(std::integral_constant<int, 0>{});
                fast_iter(std::integral_constant<int, 1>{});
                fast_iter(std::integral_constant<int, 2>{});
                fast_iter(std::integral_constant<int, 3>{});
                fast_iter(std::integral_constant<int, 4>{});
                fast_iter(std::integral_constant<int, 5>{});
                fast_iter(std::integral_constant<int, 6>{});
                fast_iter(std::integral_constant<int, 7>{});
            }
        } else {
            ptrs_valid = false;
            const int g1 = 8 * B + 8 < total ? 8 * B + 8 : total;
#pragma unroll 1
            for (int g = 8 * B; g < g1; ++g) slow_iter(g);
            ++B;
            bmod = bmod + 1 == RB ? 0 : bmod + 1;
        }
    }
    while (s_tile < n_tiles) do_store();               // the tiles that end with the block

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
