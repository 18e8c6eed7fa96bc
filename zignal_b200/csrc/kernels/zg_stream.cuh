// zignal-b200 :: streaming block evaluator skeleton for sm_100a (hand-written; the per-sample tick is
// a functor -- prebuilt for biquad cascades (zg_biquad.cuh), generated per graph for everything else).
//
// What it replaces: the caller's per-sample loop around stateful_lambda::operator()
// (reference test/benchmark.cpp:137-147 calling flowz/flowz.hpp:1225-1229), for `channels`
// independent voices at once.
//
// Design (B200):
//   * one LANE per channel: the recurrence of a voice is serial in time, voices are independent.
//   * all delay-line state and per-channel parameters stay in REGISTERS for the whole block; state
//     is read from HBM once at block start and written once at block end ([n_state][channels]).
//   * samples move HBM -> smem -> HBM with TMA (cp.async.bulk.tensor) in boxes of 32 lanes x 32
//     samples (4 KB, rows of 128 B, hardware SWIZZLE_128B).  With that swizzle a lane reads/writes
//     its own row 16 bytes at a time (LDS.128/STS.128) and every quarter-warp covers all 32 banks:
//     conflict free without padding.  Outputs overwrite the consumed inputs in place and the same
//     box is stored back with TMA, so a box costs 4 KB per wire.  A pipeline stage ("tile") is
//     `boxes` consecutive boxes in time whose loads / stores are issued back to back, so that HBM
//     sees 128*boxes contiguous bytes per channel row at a time (DRAM page locality).
//   * every WARP runs its own pipeline (own stages, own mbarriers, lane 0 issues the TMA): no
//     __syncthreads anywhere; S stages = 1 computing, 1 draining its store, S-2 loads in flight.
//   * grid = one warp per 32 channels; the host sizes CTAs so that all warps are resident at once and
//     spread evenly over the 148 SMs.
//   * sample storage in HBM is fp32 or bf16 (kIo = bytes per sample); state, parameters and all
//     arithmetic are fp32 either way.  A planar box row is always 128 bytes (32 fp32 / 64 bf16 samples),
//     a 16-byte chunk is 4 / 8 ticks; bf16 outputs are rounded to nearest even (cvt.rn.bf16x2.f32).
#pragma once

namespace zgk {

struct alignas(64) TensorMap { unsigned long long opaque[16]; };   // CUtensorMap, 128 bytes

constexpr int kMaxWires = 8;
constexpr int kTileT = 32;          // samples per fp32 box row (128 bytes)
constexpr int kTileBytes = 4096;    // 32 rows x 128 bytes
// box geometry for a sample size of `io` bytes: planar boxes are 32 channel rows of 128 bytes,
// interleaved boxes are 32 frames of 32 channels
__host__ __device__ constexpr int box_samples(bool interleaved, int io) { return interleaved ? 32 : 128 / io; }
__host__ __device__ constexpr int box_bytes(bool interleaved, int io) { return interleaved ? 1024 * io : 4096; }
constexpr int kMaxState = 64;       // state floats per channel a register-resident tick may keep
constexpr int kMaxUniform = 64;     // uniform parameters passed by value
constexpr int kMaxRingIn = 8;       // far reads of long delay lines per tick
constexpr int kMaxRingOut = 4;      // long delay lines per graph

struct StreamArgs {
    TensorMap in_map[kMaxWires];    // planar: 2-D {T, C}; interleaved: 2-D {C, T}; box 32 x 32
    TensorMap out_map[kMaxWires];
    float* state;                   // [n_state][ch_stride]
    const float* params;            // [n_params][ch_stride]
    long long ch_stride;
    long long stream_pos;           // absolute index of sample 0 of this block (dirac inputs)
    int channels;
    int n_samples;                  // samples in this block
    int boxes;                      // NB: boxes (of 32 samples) per pipeline stage
    int stages;                     // S >= 2
    int flags;                      // bit 0: in_map[1] / out_map[1] hold the 3-D whole-tile maps (K1b)
                                    // bit 1: late refill; bits 2-3: L2 hints; bit 4: L2 prefetch (see below)
                                    // bit 5 / bit 6: first / second pass of a two-pass time-segmented launch
    unsigned dirac_mask;            // synthesised input k is a dirac (else zeros)
    int state_row[kMaxState];       // row of state slot j inside `state` (prebuilt ticks use their
                                    // own slot order; generated ticks use the identity)
    float uparams[kMaxUniform];     // kUniform kernels: parameter values shared by all channels,
                                    // read straight from the constant bank (no register, no HBM)
    // Long delay lines (generated ticks only): line rows [row0, row0 + depth) of `state` are a ring, the value
    // pushed at absolute tick t sits in row t mod depth.  Ring input r (kernel input N_IN - N_RING_IN + r) reads
    // row (t - delay) mod depth; ring output w (kernel output N_OUT - N_RING_OUT + w) is stored to row t mod depth.
    int ring_in_row0[kMaxRingIn], ring_in_depth[kMaxRingIn], ring_in_delay[kMaxRingIn];
    int ring_out_row0[kMaxRingOut], ring_out_depth[kMaxRingOut];
    unsigned long long state_nowrite;   // bit j: state slot j is a window onto a ring -- loaded, never written back
    // L2 prefetch of the input rows ahead of the TMA loads (flags bit 4, planar only): every lane asks for
    // `pf_window` contiguous bytes of its own channel row `pf_dist` tiles before the tile that starts the window is
    // computed, so DRAM sees long runs per row although the shared-memory boxes stay 128 bytes wide.
    const unsigned char* in_base[kMaxWires];
    long long in_pitch_bytes;
    int pf_window;                      // bytes, a multiple of the tile's bytes per row
    int pf_dist;                        // tiles
    // Time segments (linear ticks, FAST mode; zg_runtime.cu: scan_*): with n_segs > 1 a warp is (channel group,
    // segment) instead of a channel group, and runs samples [seg * seg_len, (seg + 1) * seg_len) of its 32 channels.
    //   warm-up form (no flag): segment g > 0 starts from ZERO state `seg_warm` samples early and discards the
    //     outputs of those samples -- the host has checked that the graph forgets its state within seg_warm ticks
    //     (|A^seg_warm| below fp32 resolution), so the state at the segment's first sample is the serial one;
    //   two-pass form, any linear tick: pass 1 (flag bit 5) runs every segment but the last from zero state
    //     (segment 0 from the true state), stores no samples and leaves its final state in seg_state[g]; a small
    //     kernel turns those into the true boundary states (x <- A^seg_len x + z, zg_scan_fixup_kernel); pass 2
    //     (flag bit 6) starts segment g > 0 from seg_state[g - 1].
    // seg_len and seg_warm are multiples of the tile length; the last segment takes the ragged end of the block.
    int n_segs, seg_len, seg_warm;
    float* seg_state;                   // [n_segs][seg_state_stride]; rows laid out like `state`
    long long seg_state_stride;         // floats per segment
};

// ---- PTX wrappers -------------------------------------------------------------------------------

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap* map, int x, int y,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const TensorMap* map, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(x), "r"(y), "r"(smem_u32(src))
                 : "memory");
}
// same with an L2 eviction-priority hint (createpolicy): samples are streamed once, they need not stay in L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const TensorMap* map, int x, int y,
                                                 unsigned long long* bar, unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const TensorMap* map, int x, int y, const void* src,
                                                  unsigned long long policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(map),
                 "r"(x), "r"(y), "r"(smem_u32(src)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, unsigned bytes) {     // bytes % 16 == 0, gptr 16-byte aligned
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const TensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- sample storage <-> fp32 ---------------------------------------------------------------------
template <int kIo> struct Io;
template <> struct Io<4> {
    static constexpr int kPerChunk = 4;                       // samples per 16-byte chunk
    typedef float Elem;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    static __device__ __forceinline__ float load(const void* p) { return *reinterpret_cast<const float*>(p); }
    static __device__ __forceinline__ void store(void* p, float v) { *reinterpret_cast<float*>(p) = v; }
};
template <> struct Io<2> {                                    // bf16: the upper 16 bits of an fp32
    static constexpr int kPerChunk = 8;
    typedef unsigned short Elem;
    static __device__ __forceinline__ unsigned pack2(float lo, float hi) {
        unsigned r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
    static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
        f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
        f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    static __device__ __forceinline__ float load(const void* p) {
        return __uint_as_float((unsigned)*reinterpret_cast<const unsigned short*>(p) << 16);
    }
    static __device__ __forceinline__ void store(void* p, float v) {
        *reinterpret_cast<unsigned short*>(p) = (unsigned short)(pack2(v, 0.f) & 0xffffu);
    }
};

template <int N>
struct Arr {                       // zero-length-safe register array
    float v[N > 0 ? N : 1];
    __device__ __forceinline__ float& operator[](int i) { return v[i]; }
    __device__ __forceinline__ const float& operator[](int i) const { return v[i]; }
};

struct UniformParams {             // view of StreamArgs::uparams (kernel parameter space)
    const float* base;
    __device__ __forceinline__ float operator[](int i) const { return base[i]; }
};

__host__ __device__ constexpr int popcount_u32(unsigned v) { return v == 0 ? 0 : (int)(v & 1u) + popcount_u32(v >> 1); }

// How many of the 8 chunks of a box row are unrolled into one loop body: Tick::CHUNK_UNROLL (8, 4, 2 or 1)
// if the tick declares it, else all 8.  A 4-section EXACT biquad box is 1200 instructions (19 KB) fully
// unrolled -- with 14 warps at different places of it the instruction caches miss (ncu: no_instruction).
template <class T, class = void>
struct chunk_unroll { static constexpr int value = 8; };
template <class T>
struct chunk_unroll<T, decltype((void)T::CHUNK_UNROLL)> { static constexpr int value = T::CHUNK_UNROLL; };

// Registers a tick keeps across the block besides its delay-line state (Tick::N_EXTRA, default 0): values
// derived from the state at block start by Tick::init(s, p, e) and handed to tick(x, y, s, p, e).
template <class T, class = void>
struct extra_count { static constexpr int value = 0; };
template <class T>
struct extra_count<T, decltype((void)T::N_EXTRA)> { static constexpr int value = T::N_EXTRA; };

// Long delay lines: the last Tick::N_RING_IN inputs / N_RING_OUT outputs of a tick are not sample wires but
// reads / pushes of delay lines kept as rings in HBM (StreamArgs::ring_*).  Default: none.
template <class T, class = void>
struct ring_in_count { static constexpr int value = 0; };
template <class T>
struct ring_in_count<T, decltype((void)T::N_RING_IN)> { static constexpr int value = T::N_RING_IN; };
template <class T, class = void>
struct ring_out_count { static constexpr int value = 0; };
template <class T>
struct ring_out_count<T, decltype((void)T::N_RING_OUT)> { static constexpr int value = T::N_RING_OUT; };

// chunks (planar) / groups of four frames (interleaved) a far read is requested ahead of its use: Tick::RING_PF
template <class T, class = void>
struct ring_prefetch { static constexpr int value = 1; };
template <class T>
struct ring_prefetch<T, decltype((void)T::RING_PF)> { static constexpr int value = T::RING_PF; };

__device__ __forceinline__ int ring_mod(long long t, int depth) {     // t may be negative (before the stream began)
    int r = (int)(t % depth);
    return r < 0 ? r + depth : r;
}

// ---- the streaming loop ---------------------------------------------------------------------------
//
// Tick must provide
//   static constexpr int N_IN, N_OUT, N_STATE, N_PARAM;
//   static constexpr unsigned SYNTH_MASK;         // bit k set: input k is synthesised (dirac/zero)
//   template <class P>                            // P = Arr<N_PARAM> or UniformParams
//   static __device__ void tick(const Arr<N_IN>& x, Arr<N_OUT>& y, Arr<N_STATE>& s, const P& p);

// kSeg: the kernel understands time segments (StreamArgs::n_segs).  Only FAST-mode kernels are built with it -- a cut
// in time re-associates the arithmetic, EXACT launches never ask for one, and their code stays free of the bookkeeping.
template <class Tick, bool kInterleaved, bool kUniform, int kIo = 4, bool kSeg = false>
__device__ __forceinline__ void stream_block(const StreamArgs& a) {
    typedef Io<kIo> IO;
    constexpr int VPC = IO::kPerChunk;                         // ticks per 16-byte chunk
    constexpr int BT = box_samples(kInterleaved, kIo);        // samples per box
    constexpr int BB = box_bytes(kInterleaved, kIo);          // bytes per box
    constexpr int NRI = ring_in_count<Tick>::value, NRO = ring_out_count<Tick>::value;
    constexpr int NIT = Tick::N_IN, NOT = Tick::N_OUT;                       // all inputs / outputs of the tick
    constexpr int NI = NIT - NRI, NO = NOT - NRO;                            // sample wires (TMA, shared memory)
    constexpr int NS = Tick::N_STATE, NP = Tick::N_PARAM;
    static_assert(NRI <= kMaxRingIn && NRO <= kMaxRingOut, "too many long delay lines / far reads");
    constexpr int PF = ring_prefetch<Tick>::value;                           // >= 1
    constexpr int NT = (NI > NO ? NI : NO) > 0 ? (NI > NO ? NI : NO) : 1;   // wires per stage
    constexpr unsigned kAllIn = NI > 0 ? ((1u << NI) - 1u) : 0u;
    constexpr unsigned kBufMask = kAllIn & ~Tick::SYNTH_MASK;
    constexpr int kNumBuf = popcount_u32(kBufMask);

    extern __shared__ __align__(1024) unsigned char smem[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int S = a.stages;
    const int NB = a.boxes;
    const int tile_t = NB * BT;
    const long long gw = (long long)blockIdx.x * warps_per_cta + warp;
    // warp -> (segment, channel group): the warps of a CTA are neighbouring channel groups at the same place in time
    // (every use below goes through T_LO / T_HI / T_SEG so that a kernel built without kSeg reads a.n_samples and
    //  literal zeros exactly as it did before segments existed: ptxas allocates the hot loop differently otherwise,
    //  and the EXACT 4-section tick lost 9 % to that)
    int seg_v = 0, n_segs_v = 1, c0_v = 0, t_seg_v = 0, t_lo_v = 0, t_hi_v = 0;
    bool pass1_v = false, pass2_v = false;
    if constexpr (kSeg) {
        const int n_groups = (a.channels + 31) >> 5;
        n_segs_v = a.n_segs > 1 ? a.n_segs : 1;
        seg_v = (int)(gw / n_groups);
        if (seg_v >= n_segs_v) return;                 // warp-uniform
        pass1_v = (a.flags & 32) != 0;
        pass2_v = (a.flags & 64) != 0;
        if (pass1_v && seg_v + 1 == n_segs_v) return;  // the block's final state comes out of pass 2
        t_seg_v = n_segs_v > 1 ? seg_v * a.seg_len : 0;                             // first sample this warp answers for
        t_lo_v = seg_v > 0 && !pass1_v && !pass2_v ? t_seg_v - a.seg_warm : t_seg_v; // first sample it evaluates
        t_hi_v = seg_v + 1 < n_segs_v ? t_seg_v + a.seg_len : a.n_samples;          // end of its samples
        c0_v = (int)(gw - (long long)seg_v * n_groups) * 32;
    }
    const long long c0ll = gw * 32;
    if (!kSeg && c0ll >= a.channels) return;           // warp-uniform
    const int c0 = kSeg ? c0_v : (int)c0ll;
    const int seg = seg_v, n_segs = n_segs_v;
    const bool pass1 = pass1_v, pass2 = pass2_v;
#define T_LO (kSeg ? t_lo_v : 0)
#define T_HI (kSeg ? t_hi_v : a.n_samples)
#define T_SEG (kSeg ? t_seg_v : 0)
    const int ch = c0 + lane;
    const bool ch_ok = ch < a.channels;

    // SWIZZLE_128B works on 1024-byte atoms of the shared address: align the tile area explicitly
    // (the host reserves the slack) instead of trusting the placement of dynamic shared memory
    unsigned char* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const unsigned wire_bytes = (unsigned)NB * BB;                  // one wire of one stage
    const unsigned stage_bytes = (unsigned)NT * wire_bytes;
    unsigned char* my = tiles + (size_t)warp * S * stage_bytes;
    unsigned long long* bars =
        reinterpret_cast<unsigned long long*>(tiles + (size_t)warps_per_cta * S * stage_bytes) + warp * S;

    // ---- state and parameters into registers ----
    static_assert(NS <= kMaxState, "too much state for a register-resident tick");
    static_assert(!kUniform || NP <= kMaxUniform, "too many uniform parameters");
    Arr<NS> s;
    // segment 0 continues the stream from `state`; a later segment starts from zero (warm-up form, pass 1) or from the
    // boundary state the fix-up kernel left for it (pass 2)
    if constexpr (kSeg) {
        const float* st_in = seg == 0 ? a.state : pass2 ? a.seg_state + (long long)(seg - 1) * a.seg_state_stride : nullptr;
#pragma unroll
        for (int j = 0; j < NS; ++j) s[j] = ch_ok && st_in ? st_in[(long long)a.state_row[j] * a.ch_stride + ch] : 0.f;
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) s[j] = ch_ok ? a.state[(long long)a.state_row[j] * a.ch_stride + ch] : 0.f;
    }
    Arr<kUniform ? 0 : NP> prm_reg;
    if (!kUniform) {
#pragma unroll
        for (int j = 0; j < NP; ++j) prm_reg[j] = ch_ok ? a.params[(long long)j * a.ch_stride + ch] : 0.f;
    }
    const UniformParams prm_uni{a.uparams};
    constexpr int NE = extra_count<Tick>::value;
    Arr<NE> ex;
    if constexpr (NE > 0) {
        if constexpr (kUniform) Tick::init(s, prm_uni, ex);
        else Tick::init(s, prm_reg, ex);
    }
    auto run_tick = [&](const Arr<NIT>& x, Arr<NOT>& y) {
        if constexpr (NE > 0) {
            if constexpr (kUniform) Tick::tick(x, y, s, prm_uni, ex);
            else Tick::tick(x, y, s, prm_reg, ex);
        } else {
            if constexpr (kUniform) Tick::tick(x, y, s, prm_uni);
            else Tick::tick(x, y, s, prm_reg);
        }
    };

    if (lane == 0) {
        if (kNumBuf > 0) {
            for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
            fence_barrier_init();
        }
#pragma unroll
        for (int k = 0; k < NI; ++k)
            if (kBufMask & (1u << k)) prefetch_tmap(&a.in_map[k]);
#pragma unroll
        for (int o = 0; o < NO; ++o) prefetch_tmap(&a.out_map[o]);
    }
    __syncwarp();

    const int n_tiles = (T_HI - T_LO + tile_t - 1) / tile_t;
    // flags bit 2 / bit 3: L2 evict-first hint on the sample loads / stores
    const bool hint_loads = (a.flags & 4) != 0, hint_stores = (a.flags & 8) != 0;
    const unsigned long long policy = (a.flags & 12) ? l2_policy_evict_first() : 0ull;

    auto boxes_in_tile = [&](int t0) {                 // boxes of tile at t0 that hold samples
        const int left = (T_HI - t0 + BT - 1) / BT;
        return left < NB ? left : NB;
    };
    auto issue_load = [&](int i) {                     // lane 0 only
        if (kNumBuf == 0) return;
        const int slot = i % S;
        const int t0 = T_LO + i * tile_t;
        const int nb = boxes_in_tile(t0);
        mbar_expect_tx(&bars[slot], (unsigned)(nb * kNumBuf * BB));
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            if (!(kBufMask & (1u << k))) continue;
            unsigned char* dst = my + (size_t)slot * stage_bytes + (size_t)k * wire_bytes;
            for (int b = 0; b < nb; ++b) {
                const int cx = kInterleaved ? c0 : t0 + b * BT, cy = kInterleaved ? t0 + b * BT : c0;
                if (hint_loads) tma_load_2d_hint(dst + b * BB, &a.in_map[k], cx, cy, &bars[slot], policy);
                else tma_load_2d(dst + b * BB, &a.in_map[k], cx, cy, &bars[slot]);
            }
        }
    };

    if (lane == 0) {
        const int pre = n_tiles < S - 1 ? n_tiles : S - 1;
        for (int i = 0; i < pre; ++i) issue_load(i);
    }
    // The slot that tile i-1 used is refilled (tile i+S-1) as soon as the first box of tile i is done:
    // by then the store of tile i-1 has long read its shared memory, and the load has the rest of the
    // tile's arithmetic to land -- a warp never waits a full HBM round trip at a tile boundary.
    // (flags bit 1: refill only after the store of tile i has been issued, the first version.)
    const bool early_refill = (a.flags & 2) == 0;

    // ---- long delay lines: ring rows of `state`, one coalesced 4-byte access per lane (128 bytes per warp) ----
    int rin_idx[NRI > 0 ? NRI : 1];                    // next row to load, per ring input
    int rout_idx[NRO > 0 ? NRO : 1];                   // next row to store, per ring output
    auto ring_seek = [&](long long t_abs) {
#pragma unroll
        for (int r = 0; r < NRI; ++r) rin_idx[r] = ring_mod(t_abs - a.ring_in_delay[r], a.ring_in_depth[r]);
#pragma unroll
        for (int w = 0; w < NRO; ++w) rout_idx[w] = ring_mod(t_abs, a.ring_out_depth[w]);
    };
    auto ring_load = [&](int r) {
        const float v = a.state[(long long)(a.ring_in_row0[r] + rin_idx[r]) * a.ch_stride + ch];
        rin_idx[r] = rin_idx[r] + 1 == a.ring_in_depth[r] ? 0 : rin_idx[r] + 1;
        return v;
    };
    auto ring_store = [&](int w, float v) {
        a.state[(long long)(a.ring_out_row0[w] + rout_idx[w]) * a.ch_stride + ch] = v;
        rout_idx[w] = rout_idx[w] + 1 == a.ring_out_depth[w] ? 0 : rout_idx[w] + 1;
    };

    for (int i = 0; i < n_tiles; ++i) {
        const int slot = i % S;
        const int t0 = T_LO + i * tile_t;
        const int nb = boxes_in_tile(t0);
        unsigned char* stage = my + (size_t)slot * stage_bytes;

        if (!kInterleaved && kNumBuf > 0 && (a.flags & 16) && ch_ok) {
            const long long off = kSeg ? ((long long)t0 + (long long)a.pf_dist * tile_t) * kIo
                                       : (long long)(i + a.pf_dist) * tile_t * kIo;          // byte offset inside the row
            const long long row_bytes = ((long long)a.n_samples * kIo) & ~15ll;
            if (off % a.pf_window == 0 && off < row_bytes) {
                const unsigned sz = (unsigned)(row_bytes - off < a.pf_window ? row_bytes - off : a.pf_window);
#pragma unroll
                for (int k = 0; k < NI; ++k)
                    if (kBufMask & (1u << k)) l2_prefetch_bulk(a.in_base[k] + (long long)ch * a.in_pitch_bytes + off, sz);
            }
        }
        if (kNumBuf > 0) mbar_wait(&bars[slot], (unsigned)((i / S) & 1));

        if (ch_ok) {
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const int tb0 = t0 + b * BT;
                const int n_valid = T_HI - tb0 < BT ? T_HI - tb0 : BT;
                unsigned char* base = stage + b * BB;              // wire k of this box: base + k * wire_bytes
                if (early_refill && lane == 0 && b == 1 && i + S - 1 < n_tiles) {
                    tma_wait_read<0>();                            // the store of tile i-1 (the only one pending)
                    issue_load(i + S - 1);
                }
                // absolute stream index of the first sample of the box (dirac synthesis).  A dirac is 1.0 at
                // stream position 0 only: the one box that holds it takes the sample-by-sample path below;
                // everywhere else a synthesised input is the constant 0 (no per-sample compare in the hot loop)
                const long long t_abs0 = a.stream_pos + tb0;
                const bool full_box = n_valid == BT && !(Tick::SYNTH_MASK != 0 && a.dirac_mask != 0 && t_abs0 == 0);

                if (!kInterleaved) {
                    // row `lane`, 16-byte chunk j lives at chunk (j ^ (lane & 7)) of the row (SWIZZLE_128B)
                    const unsigned row = (unsigned)lane * 128u;
                    const unsigned sw = (unsigned)(lane & 7);
                    if (full_box) {
                        // the chunk of step j+1 is loaded before the ticks of step j: the LDS latency
                        // overlaps arithmetic instead of stalling the warp (3-4 warps per scheduler)
                        uint4 xn[NI > 0 ? NI : 1];
#pragma unroll
                        for (int k = 0; k < NI; ++k)
                            if (kBufMask & (1u << k))
                                xn[k] = *reinterpret_cast<const uint4*>(base + k * wire_bytes + row + (sw << 4));
                        // far reads of long delay lines: the rows of chunk j+1 are requested before the ticks of
                        // chunk j (they were stored at least two chunks ago: delay >= 2 * VPC, zg_ir.hpp)
                        float rn[PF][NRI > 0 ? NRI : 1][VPC];            // rows of chunks j .. j+PF-1, in flight
                        ring_seek(t_abs0);
#pragma unroll
                        for (int f = 0; f < PF; ++f)
#pragma unroll
                            for (int r = 0; r < NRI; ++r)
#pragma unroll
                                for (int q = 0; q < VPC; ++q) rn[f][r][q] = ring_load(r);
                        constexpr int CU = chunk_unroll<Tick>::value;
                        static_assert(CU == 8 || CU == 4 || CU == 2 || CU == 1, "CHUNK_UNROLL must divide 8");
#pragma unroll 1
                        for (int h = 0; h < 8; h += CU)
#pragma unroll
                        for (int jj = 0; jj < CU; ++jj) {
                            const int j = h + jj;
                            const unsigned off = row + (((unsigned)j ^ sw) << 4);
                            float xv[NIT > 0 ? NIT : 1][VPC];
                            float yv[NOT > 0 ? NOT : 1][VPC];
#pragma unroll
                            for (int k = 0; k < NI; ++k) {
                                if (kBufMask & (1u << k)) {
                                    IO::unpack(xn[k], xv[k]);
                                    if (j < 7)
                                        xn[k] = *reinterpret_cast<const uint4*>(base + k * wire_bytes + row +
                                                                                (((unsigned)(j + 1) ^ sw) << 4));
                                } else {
#pragma unroll
                                    for (int q = 0; q < VPC; ++q) xv[k][q] = 0.f;
                                }
                            }
#pragma unroll
                            for (int r = 0; r < NRI; ++r) {
#pragma unroll
                                for (int q = 0; q < VPC; ++q) xv[NI + r][q] = rn[0][r][q];
#pragma unroll
                                for (int f = 0; f + 1 < PF; ++f)
#pragma unroll
                                    for (int q = 0; q < VPC; ++q) rn[f][r][q] = rn[f + 1][r][q];
                                if (j + PF < 8) {
#pragma unroll
                                    for (int q = 0; q < VPC; ++q) rn[PF - 1][r][q] = ring_load(r);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < VPC; ++q) {
                                Arr<NIT> x;
                                Arr<NOT> y;
#pragma unroll
                                for (int k = 0; k < NIT; ++k) x[k] = xv[k][q];
                                run_tick(x, y);
#pragma unroll
                                for (int o = 0; o < NOT; ++o) yv[o][q] = y[o];
                            }
#pragma unroll
                            for (int o = 0; o < NO; ++o)
                                *reinterpret_cast<uint4*>(base + o * wire_bytes + off) = IO::pack(yv[o]);
#pragma unroll
                            for (int w = 0; w < NRO; ++w)
#pragma unroll
                                for (int q = 0; q < VPC; ++q) ring_store(w, yv[NO + w][q]);
                        }
                    } else {
                        // last, partial box of the block: TMA zero-filled the tail on load and clips it
                        // on store; only the state has to be protected
                        ring_seek(t_abs0);
                        for (int t = 0; t < n_valid; ++t) {
                            const unsigned off = row + ((((unsigned)t / VPC) ^ sw) << 4) + ((unsigned)t % VPC) * kIo;
                            Arr<NIT> x;
                            Arr<NOT> y;
#pragma unroll
                            for (int k = 0; k < NI; ++k) {
                                if (kBufMask & (1u << k)) x[k] = IO::load(base + k * wire_bytes + off);
                                else x[k] = (((a.dirac_mask >> k) & 1u) && t_abs0 + t == 0) ? 1.f : 0.f;
                            }
#pragma unroll
                            for (int r = 0; r < NRI; ++r) x[NI + r] = ring_load(r);
                            run_tick(x, y);
#pragma unroll
                            for (int o = 0; o < NO; ++o) IO::store(base + o * wire_bytes + off, y[o]);
#pragma unroll
                            for (int w = 0; w < NRO; ++w) ring_store(w, y[NO + w]);
                        }
                    }
                } else {
                    // interleaved frames: box is [32 samples][32 channels], lane = channel column
                    typename IO::Elem* box = reinterpret_cast<typename IO::Elem*>(base);
                    const unsigned wire_f = wire_bytes / kIo;
                    if (full_box) {
                        // groups of four frames; the ring rows of group g+1 are requested before the ticks of group g
                        float rn[PF][NRI > 0 ? NRI : 1][4];
                        ring_seek(t_abs0);
#pragma unroll
                        for (int f = 0; f < PF; ++f)
#pragma unroll
                            for (int r = 0; r < NRI; ++r)
#pragma unroll
                                for (int q = 0; q < 4; ++q) rn[f][r][q] = ring_load(r);
#pragma unroll 2
                        for (int gq = 0; gq < BT / 4; ++gq) {
                            float rc[NRI > 0 ? NRI : 1][4];
#pragma unroll
                            for (int r = 0; r < NRI; ++r) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) rc[r][q] = rn[0][r][q];
#pragma unroll
                                for (int f = 0; f + 1 < PF; ++f)
#pragma unroll
                                    for (int q = 0; q < 4; ++q) rn[f][r][q] = rn[f + 1][r][q];
                                if (gq + PF < BT / 4) {
#pragma unroll
                                    for (int q = 0; q < 4; ++q) rn[PF - 1][r][q] = ring_load(r);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int t = 4 * gq + q;
                                Arr<NIT> x;
                                Arr<NOT> y;
#pragma unroll
                                for (int k = 0; k < NI; ++k) {
                                    if (kBufMask & (1u << k)) x[k] = IO::load(&box[k * wire_f + t * 32 + lane]);
                                    else x[k] = 0.f;
                                }
#pragma unroll
                                for (int r = 0; r < NRI; ++r) x[NI + r] = rc[r][q];
                                run_tick(x, y);
#pragma unroll
                                for (int o = 0; o < NO; ++o) IO::store(&box[o * wire_f + t * 32 + lane], y[o]);
#pragma unroll
                                for (int w = 0; w < NRO; ++w) ring_store(w, y[NO + w]);
                            }
                        }
                    } else {
                        ring_seek(t_abs0);
                        for (int t = 0; t < n_valid; ++t) {
                            Arr<NIT> x;
                            Arr<NOT> y;
#pragma unroll
                            for (int k = 0; k < NI; ++k) {
                                if (kBufMask & (1u << k)) x[k] = IO::load(&box[k * wire_f + t * 32 + lane]);
                                else x[k] = (((a.dirac_mask >> k) & 1u) && t_abs0 + t == 0) ? 1.f : 0.f;
                            }
#pragma unroll
                            for (int r = 0; r < NRI; ++r) x[NI + r] = ring_load(r);
                            run_tick(x, y);
#pragma unroll
                            for (int o = 0; o < NO; ++o) IO::store(&box[o * wire_f + t * 32 + lane], y[o]);
#pragma unroll
                            for (int w = 0; w < NRO; ++w) ring_store(w, y[NO + w]);
                        }
                    }
                }
            }
        }

        fence_proxy_async();                           // generic-proxy writes -> visible to TMA
        __syncwarp();
        if (lane == 0) {
            // warm-up tiles and the whole of pass 1 produce state, not samples: nothing is stored (the commit
            // group stays, empty, so that the refill logic below counts the same groups)
            const bool discard = kSeg && (pass1 || t0 < T_SEG);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                for (int b = 0; b < (discard ? 0 : nb); ++b) {
                    const unsigned char* src = stage + (size_t)o * wire_bytes + b * BB;
                    const int cx = kInterleaved ? c0 : t0 + b * BT, cy = kInterleaved ? t0 + b * BT : c0;
                    if (hint_stores) tma_store_2d_hint(&a.out_map[o], cx, cy, src, policy);
                    else tma_store_2d(&a.out_map[o], cx, cy, src);
                }
            }
            tma_commit();
            // refill the slot that tile i-1 used: its store (committed one iteration ago) must have
            // finished reading smem
            const int nxt = i + S - 1;
            if (nxt < n_tiles && !(early_refill && nb > 1)) {
                tma_wait_read<1>();
                issue_load(nxt);
            }
        }
    }

    if (lane == 0) tma_wait_all<0>();                  // smem must outlive the last stores

    // ---- state back to HBM: the last segment ends the block; pass 1 leaves every segment's final state for the fix-up ----
    float* st_out = a.state;
    if constexpr (kSeg) st_out = pass1 ? a.seg_state + (long long)seg * a.seg_state_stride : seg + 1 == n_segs ? a.state : nullptr;
    if (ch_ok && (!kSeg || st_out)) {
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (!((a.state_nowrite >> j) & 1ull)) st_out[(long long)a.state_row[j] * a.ch_stride + ch] = s[j];
    }
}

#undef T_LO
#undef T_HI
#undef T_SEG

}  // namespace zgk
