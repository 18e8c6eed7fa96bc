// zignal-b200 :: device half of the C ABI -- plans, kernel selection, NVRTC specialisation, launches.
//
// There is no CPU fallback in this file: every zg_process() is a kernel launch on the plan's B200,
// and every failure (no device, no driver, NVRTC missing, bad architecture) is returned as
// ZG_ERR_CUDA with the reason in zg_last_error().
//
// Kernel selection at plan time:
//   K1  tick program is a cascade of direct-form-1 biquads (zg_match.cpp)  -> prebuilt kernel
//       zg_biquad_df1<SECTIONS, exact, interleaved, uniform> compiled by nvcc into this library; per launch, by shape:
//       K1s (many channels: sections spread over the warps of a persistent CTA, kernels/zg_biquad_split.cuh),
//       K1b (few channels: sections across the lanes of a warp) or K1 cut in time (FAST, linear ticks);
//   K2  anything else -> the same hand-written streaming skeleton (kernels/zg_stream.cuh) with the
//       straight-line tick body generated from the graph (zg_codegen.cpp), compiled once per plan
//       with NVRTC for sm_100a.
#include <cuda.h>            // driver API *types* only; entry points come from cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a profiler has injected its library
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "kernels/zg_biquad.cuh"
#include "kernels/zg_biquad_lanes.cuh"
#include "kernels/zg_biquad_split.cuh"
#include "kernels/zg_fir.cuh"
#include "kernels/zg_fir_tc.cuh"
#include "zg_internal.hpp"

extern const char* const zg_stream_cuh_source;   // kernels/zg_stream.cuh as text (generated at build time)

using namespace zg;

namespace {

// ------------------------------------------------------------------------------------------------
// driver API + NVRTC, resolved lazily (the library must load on a machine without libcuda)
// ------------------------------------------------------------------------------------------------

struct Driver {
    decltype(&cuTensorMapEncodeTiled) tensorMapEncodeTiled = nullptr;
    decltype(&cuModuleLoadData) moduleLoadData = nullptr;
    decltype(&cuModuleUnload) moduleUnload = nullptr;
    decltype(&cuModuleGetFunction) moduleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) funcSetAttribute = nullptr;
    decltype(&cuFuncGetAttribute) funcGetAttribute = nullptr;
    decltype(&cuLaunchKernel) launchKernel = nullptr;
    decltype(&cuGetErrorString) getErrorString = nullptr;
    bool ok = false;
    std::string why;
};

template <class F>
bool drv_sym(const char* name, F& f, std::string& why) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        why = std::string("CUDA driver entry point ") + name + " unavailable: " +
              (e != cudaSuccess ? cudaGetErrorString(e) : "symbol not found");
        (void)cudaGetLastError();
        return false;
    }
    f = reinterpret_cast<F>(p);
    return true;
}

Driver& driver() {
    static Driver d = [] {
        Driver r;
        r.ok = drv_sym("cuTensorMapEncodeTiled", r.tensorMapEncodeTiled, r.why) &&
               drv_sym("cuModuleLoadData", r.moduleLoadData, r.why) &&
               drv_sym("cuModuleUnload", r.moduleUnload, r.why) &&
               drv_sym("cuModuleGetFunction", r.moduleGetFunction, r.why) &&
               drv_sym("cuFuncSetAttribute", r.funcSetAttribute, r.why) &&
               drv_sym("cuFuncGetAttribute", r.funcGetAttribute, r.why) &&
               drv_sym("cuLaunchKernel", r.launchKernel, r.why) &&
               drv_sym("cuGetErrorString", r.getErrorString, r.why);
        return r;
    }();
    return d;
}

std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (driver().getErrorString) driver().getErrorString(r, &s);
    return s ? s : ("CUresult " + std::to_string((int)r));
}

// NVRTC through dlopen: only plans that need a generated kernel touch it.
struct Nvrtc {
    void* lib = nullptr;
    int (*createProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*destroyProgram)(void**) = nullptr;
    int (*compileProgram)(void*, int, const char* const*) = nullptr;
    int (*getCUBINSize)(void*, size_t*) = nullptr;
    int (*getCUBIN)(void*, char*) = nullptr;
    int (*getProgramLogSize)(void*, size_t*) = nullptr;
    int (*getProgramLog)(void*, char*) = nullptr;
    const char* (*getErrorString)(int) = nullptr;
    int (*version)(int*, int*) = nullptr;
    bool ok = false;
    std::string why;
};

Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char* nm : names) {
            r.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (r.lib) break;
        }
        if (!r.lib) {
            r.why = std::string("cannot load libnvrtc.so.12 (needed to specialise the kernel for this graph): ") + dlerror();
            return r;
        }
        auto sym = [&](const char* s) { return dlsym(r.lib, s); };
        r.createProgram = (decltype(r.createProgram))sym("nvrtcCreateProgram");
        r.destroyProgram = (decltype(r.destroyProgram))sym("nvrtcDestroyProgram");
        r.compileProgram = (decltype(r.compileProgram))sym("nvrtcCompileProgram");
        r.getCUBINSize = (decltype(r.getCUBINSize))sym("nvrtcGetCUBINSize");
        r.getCUBIN = (decltype(r.getCUBIN))sym("nvrtcGetCUBIN");
        r.getProgramLogSize = (decltype(r.getProgramLogSize))sym("nvrtcGetProgramLogSize");
        r.getProgramLog = (decltype(r.getProgramLog))sym("nvrtcGetProgramLog");
        r.getErrorString = (decltype(r.getErrorString))sym("nvrtcGetErrorString");
        r.version = (decltype(r.version))sym("nvrtcVersion");
        r.ok = r.createProgram && r.destroyProgram && r.compileProgram && r.getCUBINSize && r.getCUBIN &&
               r.getProgramLogSize && r.getProgramLog && r.getErrorString;
        if (!r.ok) r.why = "libnvrtc is missing expected symbols";
        return r;
    }();
    return n;
}

// NVTX ranges around the host-side phases of the path (tracing): plan creation, NVRTC
// specialisation, one zg_process launch, zg_process_host and each of its row chunks.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Calls on a plan run on the plan's device and leave the caller's current device as they found it (a single-process
// multi-GPU caller -- torch included -- must not find itself on another GPU after a call).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; (void)cudaGetLastError(); }
        if (prev != device) err = cudaSetDevice(device);
        else prev = -1;                                   // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ZG_ON_DEVICE(dev)                 \
    DeviceGuard device_guard_(dev);       \
    if (device_guard_.err != cudaSuccess) return cuda_fail(device_guard_.err, "cudaSetDevice")

int cuda_fail(cudaError_t e, const char* what);

int cuda_fail(cudaError_t e, const char* what) {
    (void)cudaGetLastError();
    return fail(ZG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define ZG_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// prebuilt kernels (K1)
// ------------------------------------------------------------------------------------------------

template <class Tick, bool kInterleaved, bool kUniform, bool kSeg = false>
__global__ void __launch_bounds__(512, 1) zg_stream_kernel(const __grid_constant__ zgk::StreamArgs a) {
    zgk::stream_block<Tick, kInterleaved, kUniform, 4, kSeg>(a);
}

template <int S, bool kExact, bool kUniform>
__global__ void __launch_bounds__(512, 1) zg_biquad_lanes_kernel(const __grid_constant__ zgk::StreamArgs a) {
    zgk::biquad_lanes_block<S, kExact, kUniform>(a);
}

template <int S, int SPW, bool kExact, bool kSym, bool kUniform, bool kAllArrive = false, int kHB = 1>
__global__ void __launch_bounds__(512, 1) zg_biquad_split_kernel(const __grid_constant__ zgk::SplitArgs a) {
    zgk::biquad_split_block<S, SPW, kExact, kSym, kUniform, kAllArrive, kHB>(a);
}

template <bool kExact, bool kInterleaved>
__global__ void __launch_bounds__(512, 1) zg_fir_kernel(const __grid_constant__ zgk::FirArgs a) {
    zgk::fir_block<kExact, kInterleaved>(a);
}

// Boundary fix-up of the two-pass time-segmented form (kernels/zg_stream.cuh, StreamArgs::n_segs): pass 1 left in
// seg[g] the state segment g ends with when it starts from zero (g = 0: from the true state, so seg[0] is already
// the state segment 1 starts from).  One thread per channel walks the boundaries in order,
//      x_{g+1} = A^L x_g + z_g          (the tick is linear: flowz/flowz.hpp:1031-1074 evaluated L times),
// and overwrites seg[g] with the true state at the start of segment g + 1.  A^L is computed on the host in float64
// from the tick program (zg_scan.cpp) and shared by all channels, or one matrix per channel ([n*n][ch_stride]).
__global__ void zg_scan_fixup_kernel(float* seg, long long seg_stride, long long ch_stride, int channels, int n,
                                     int n_bounds, const float* __restrict__ AL, int per_channel) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    float x[zgk::kMaxState], y[zgk::kMaxState];
    for (int i = 0; i < n; ++i) x[i] = seg[(long long)i * ch_stride + c];
    for (int g = 1; g < n_bounds; ++g) {
        float* z = seg + (long long)g * seg_stride;
        for (int i = 0; i < n; ++i) {
            float acc = z[(long long)i * ch_stride + c];
            for (int j = 0; j < n; ++j) {
                const float a = per_channel ? AL[(long long)(i * n + j) * ch_stride + c] : AL[i * n + j];
                acc = fmaf(a, x[j], acc);
            }
            y[i] = acc;
        }
        for (int i = 0; i < n; ++i) {
            x[i] = y[i];
            z[(long long)i * ch_stride + c] = y[i];
        }
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel FUNCTION (per device), which every plan of the
// process shares: remember the largest value set so far per (function, device) and only ever raise it.
int raise_max_smem(const void* fn, int device, int smem) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> set_so_far;
    std::lock_guard<std::mutex> lock(mu);
    int& cur = set_so_far[{fn, device}];
    if (smem <= cur) return ZG_OK;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ZG_ERR_CUDA, std::string("cudaFuncSetAttribute(max dynamic smem): ") + cudaGetErrorString(e));
    }
    cur = smem;
    return ZG_OK;
}

__global__ void __launch_bounds__(zgk::kTcThreads, 1) zg_fir_tc_kernel(const __grid_constant__ zgk::FirTcArgs a) {
    zgk::fir_tc_block(a);
}

const void* fir_kernel_for(bool exact, bool interleaved) {
    if (exact) return interleaved ? (const void*)zg_fir_kernel<true, true> : (const void*)zg_fir_kernel<true, false>;
    return interleaved ? (const void*)zg_fir_kernel<false, true> : (const void*)zg_fir_kernel<false, false>;
}

using KernelPtr = void (*)(zgk::StreamArgs);

KernelPtr biquad_lanes_kernel_for(int sections, bool exact, bool uniform) {
#define ZG_PICK(S, E, U) \
    if (sections == S && exact == E && uniform == U) return (KernelPtr)zg_biquad_lanes_kernel<S, E, U>;
    ZG_PICK(2, false, false) ZG_PICK(2, false, true) ZG_PICK(2, true, false) ZG_PICK(2, true, true)
    ZG_PICK(4, false, false) ZG_PICK(4, false, true) ZG_PICK(4, true, false) ZG_PICK(4, true, true)
#undef ZG_PICK
    return nullptr;
}

int tune_env(const char* name);          // ZG_TUNE_* overrides, below

// K1s (kernels/zg_biquad_split.cuh): the sections of a channel group spread over `sections / spw` warps
using SplitKernelPtr = void (*)(zgk::SplitArgs);
SplitKernelPtr biquad_split_kernel_for(int sections, int spw, bool exact, bool sym, bool uniform, int hand_boxes = 1) {
    // few channels, one group per SM: several boxes per hand-over (kernels/zg_biquad_split.cuh kHB); 4 sections only, no
    // product reuse (a lone warp waits for its recurrence, not for issue slots)
    if (hand_boxes > 1) {
        if (sections != 4 || spw != 1) return nullptr;
#define ZG_PICK3(H, E) \
    if (hand_boxes == H && exact == E) \
        return uniform ? (SplitKernelPtr)zg_biquad_split_kernel<4, 1, E, false, true, false, H> : (SplitKernelPtr)zg_biquad_split_kernel<4, 1, E, false, false, false, H>;
        if (tune_env("ZG_TUNE_SPLIT_ARRIVE") == 1 && hand_boxes == 4)      // the form the race checker can follow
            return exact ? (uniform ? (SplitKernelPtr)zg_biquad_split_kernel<4, 1, true, false, true, true, 4>
                                    : (SplitKernelPtr)zg_biquad_split_kernel<4, 1, true, false, false, true, 4>)
                         : (uniform ? (SplitKernelPtr)zg_biquad_split_kernel<4, 1, false, false, true, true, 4>
                                    : (SplitKernelPtr)zg_biquad_split_kernel<4, 1, false, false, false, true, 4>);
        ZG_PICK3(2, false) ZG_PICK3(2, true) ZG_PICK3(4, false) ZG_PICK3(4, true)
#undef ZG_PICK3
        return nullptr;
    }
    // the form the race checker can follow (kernels/zg_biquad_split.cuh kAllArrive): the 4-section kernel only
    if (tune_env("ZG_TUNE_SPLIT_ARRIVE") == 1 && sections == 4 && spw == 1) {
#define ZG_PICK3(E, Y) \
    if (exact == E && sym == Y) \
        return uniform ? (SplitKernelPtr)zg_biquad_split_kernel<4, 1, E, Y, true, true> : (SplitKernelPtr)zg_biquad_split_kernel<4, 1, E, Y, false, true>;
        ZG_PICK3(false, false) ZG_PICK3(true, false) ZG_PICK3(true, true)
#undef ZG_PICK3
    }
#define ZG_PICK3(S, W, E, Y) \
    if (sections == S && spw == W && exact == E && sym == Y) \
        return uniform ? (SplitKernelPtr)zg_biquad_split_kernel<S, W, E, Y, true> : (SplitKernelPtr)zg_biquad_split_kernel<S, W, E, Y, false>;
#define ZG_PICK(S, W) ZG_PICK3(S, W, false, false) ZG_PICK3(S, W, true, false) ZG_PICK3(S, W, true, true)
    ZG_PICK(2, 1) ZG_PICK(3, 1) ZG_PICK(4, 1) ZG_PICK(5, 1) ZG_PICK(6, 1) ZG_PICK(7, 1) ZG_PICK(8, 1) ZG_PICK(4, 2) ZG_PICK(6, 2) ZG_PICK(8, 2)
#undef ZG_PICK
#undef ZG_PICK3
    return nullptr;
}

template <int S>
KernelPtr biquad_kernel(bool exact, bool interleaved, bool uniform, bool sym, bool seg) {
    // FAST kernels exist with and without time segments (kernels/zg_stream.cuh kSeg): the bookkeeping costs the
    // unsegmented launch 3.5 % on the north-star shape, so it only runs when a launch is actually cut in time
    if (seg && !exact) {
#define ZG_PICK(I, U) \
    if (interleaved == I && uniform == U) return (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, false>, I, U, true>;
        ZG_PICK(false, false) ZG_PICK(false, true) ZG_PICK(true, false) ZG_PICK(true, true)
#undef ZG_PICK
    }
    // b0 == b2 in every section (of every channel), EXACT: the product-reusing tick (kernels/zg_biquad.cuh)
    if (sym && exact) {
        if (uniform)
            return interleaved ? (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, true, true>, true, true>
                               : (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, true, true>, false, true>;
        return interleaved ? (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, true, true>, true, false>
                           : (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, true, true>, false, false>;
    }
#define ZG_PICK(E, I, U) \
    if (exact == E && interleaved == I && uniform == U) \
        return (KernelPtr)zg_stream_kernel<zgk::BiquadDf1Cascade<S, E>, I, U>;
    ZG_PICK(false, false, false) ZG_PICK(false, false, true) ZG_PICK(false, true, false) ZG_PICK(false, true, true)
    ZG_PICK(true, false, false) ZG_PICK(true, false, true) ZG_PICK(true, true, false) ZG_PICK(true, true, true)
#undef ZG_PICK
    return nullptr;
}

KernelPtr biquad_kernel_for(int sections, bool exact, bool interleaved, bool uniform, bool sym, bool seg) {
    switch (sections) {
        case 1: return biquad_kernel<1>(exact, interleaved, uniform, sym, seg);
        case 2: return biquad_kernel<2>(exact, interleaved, uniform, sym, seg);
        case 3: return biquad_kernel<3>(exact, interleaved, uniform, sym, seg);
        case 4: return biquad_kernel<4>(exact, interleaved, uniform, sym, seg);
        case 5: return biquad_kernel<5>(exact, interleaved, uniform, sym, seg);
        case 6: return biquad_kernel<6>(exact, interleaved, uniform, sym, seg);
        case 7: return biquad_kernel<7>(exact, interleaved, uniform, sym, seg);
        case 8: return biquad_kernel<8>(exact, interleaved, uniform, sym, seg);
    }
    return nullptr;
}

struct Variant {            // one compiled kernel of a plan
    bool ready = false;
    KernelPtr prebuilt = nullptr;
    CUmodule module = nullptr;
    CUfunction function = nullptr;
    int regs = 0;
    int max_smem_set = 0;
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// the plan
// ------------------------------------------------------------------------------------------------

struct zg_plan {
    Ir ir;                                  // private copy: the plan outlives nothing it points to
    zg_plan_opts opts{};
    int64_t C = 0, ch_stride = 0;
    bool exact = false, interleaved = false;
    int io = 4;                             // bytes per sample in HBM: 4 = fp32, 2 = bf16
    int tick_ops = 0;                       // floating-point instructions per tick (estimate)
    bool light_tick = false;                // < 3.5 such instructions per byte of sample traffic: HBM-bound
                                            // (refill and geometry policy); else bound by instruction issue
    unsigned synth_mask = 0, dirac_mask = 0;
    int n_buf_in = 0;

    bool is_biquad = false;
    int lanes = 1;                          // lanes per channel (K1b when > 1)
    BiquadMatch bq;
    int kernel_n_state = 0, kernel_n_param = 0;   // as the kernel sees them
    std::vector<int> state_row;                   // kernel slot -> row of d_state

    Ir kir;                                 // K2: the kernel-side tick program (long delay lines split off, zg_ir.hpp)
    RingPlan ring;                          //     and where its extra inputs / outputs / window slots live
    int ring_pf = 1;                        //     chunks a far read is requested ahead of use
    bool fir_tc = false;                    // K3t: the FIR on tensor cores (kernels/zg_fir_tc.cuh): FAST, planar, <= 256 taps
    bool is_fir = false;                    // K3: dense FIR (kernels/zg_fir.cuh)
    FirMatch fir;
    float* d_taps = nullptr;                // [n_taps]
    float* d_state_alt = nullptr;           // K3 ping-pongs the delay line between two buffers
    int fir_regs = 0;
    int last_grid = 0;

    Variant variant[12];                    // [0] per-channel parameters, [1] uniform, [2], [3] the same with symmetric biquads;
                                            // [4], [5]: the lane-per-channel kernel of a K1b plan; + 6: built with time segments
    int lanes_now = 1;                      // lanes per channel of the launch being prepared
    bool split_now = false;                 // the last launch ran K1s (sections spread over the warps of a group)
    int split_regs = 0, split_spw = 0, split_hand_boxes = 1;
    const void* split_fn = nullptr;
    unsigned long long* d_split_ctl = nullptr;     // K1s: tickets drawn / CTAs finished / launches finished (kept by the kernel)
    unsigned long long* d_split_flags = nullptr;   // [channel groups][warps per group]: epoch of the row's head piece
    float* d_split_carry = nullptr;                // [warps per group * state per warp][ch_stride]: delay lines of a cut row
    int split_wpg_alloc = 0, split_segs_alloc = 0;
    bool seg_now = false;                   // the launch being prepared is cut in time

    // Time segments for few, long channels (FAST mode, linear ticks; kernels/zg_stream.cuh StreamArgs::n_segs)
    int linearity = ZG_NONLINEAR;
    bool scan_ok = false;                   // graph and options allow a time-segmented launch
    bool scan_dirty = true;                 // parameters changed since the analysis below
    int scan_warm = 0;                      // ticks after which every channel has forgotten its state (|A^K| <= 2^-30);
                                            // 0: some channel never does
    float* d_seg_state = nullptr;           // [segments][n_state][ch_stride]: boundary states of the two-pass form
    size_t seg_state_floats = 0;
    float* d_AL = nullptr;                  // A^L: [n*n] shared, or [n*n][ch_stride] per channel
    size_t AL_floats = 0;
    int AL_len = 0;                         // the L it was computed for (0: stale)
    bool AL_per_channel = false;
    int last_segs = 1, last_seg_mode = 0, last_seg_len = 0, last_seg_warm = 0;
    bool sym_now = false;                   // K1: every section of every channel has b0 == b2 (bit for bit)
    std::string kernel_name;

    float* d_state = nullptr;               // [n_state][ch_stride]
    float* d_params = nullptr;              // [kernel_n_param][ch_stride]
    std::vector<std::vector<float>> h_params;   // per graph parameter: size 1 (scalar) or C
    std::vector<char> param_on_device;          // parameter k was last set from device memory (zg_param_set_device)
    float* d_user_params = nullptr;             // [n_params][ch_stride]: the device-side values of those
    bool params_dirty = true;
    bool uniform_now = true;
    float uparams[zgk::kMaxUniform] = {};

    int64_t stream_pos = 0;
    int launches = 0;
    int last_smem = 0, last_threads = 0, last_stages = 0, last_boxes = 0;
    int sm_count = 148;
    int max_smem_optin = 227 * 1024;

    // staging for zg_process_host
    unsigned char* d_stage = nullptr;
    size_t d_stage_bytes = 0;
    cudaStream_t own_stream = nullptr;      // kernels of zg_process_host
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> events;        // [2 * chunk]: input landed, kernel done
    int last_host_chunks = 0;

    ~zg_plan() {
        int prev_device = -1;
        if (cudaGetDevice(&prev_device) != cudaSuccess) prev_device = -1;
        if (opts.device >= 0) cudaSetDevice(opts.device);
        for (auto& v : variant)
            if (v.module && driver().moduleUnload) driver().moduleUnload(v.module);
        if (d_state) cudaFree(d_state);
        if (d_state_alt) cudaFree(d_state_alt);
        if (d_taps) cudaFree(d_taps);
        if (d_params) cudaFree(d_params);
        if (d_user_params) cudaFree(d_user_params);
        if (d_stage) cudaFree(d_stage);
        if (d_seg_state) cudaFree(d_seg_state);
        if (d_AL) cudaFree(d_AL);
        if (d_split_ctl) cudaFree(d_split_ctl);
        if (d_split_flags) cudaFree(d_split_flags);
        if (d_split_carry) cudaFree(d_split_carry);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (h2d_stream) cudaStreamDestroy(h2d_stream);
        if (d2h_stream) cudaStreamDestroy(d2h_stream);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        if (prev_device >= 0) cudaSetDevice(prev_device);
    }
};

namespace {

// ---- NVRTC specialisation (K2) -----------------------------------------------------------------

constexpr int kRegLineDepth = 16;           // delay lines up to this depth stay in registers whatever their reads

// Far reads are requested `pf` chunks (of `ticks per 16 bytes`; groups of four frames when interleaved) ahead of
// use and must have been stored before the request: a read is far from delay (pf + 1) * chunk on.
int ring_chunk(bool interleaved, int io) { return interleaved ? 4 : 16 / io; }

// The kernel-side program of a graph: two chunks of prefetch when the in-flight rows fit in <= 64 registers,
// else one (more reads become far with the shorter distance; the caller checks the limits).
Ir split_for_kernel(const Ir& ir, bool interleaved, int io, RingPlan& ring, int& pf) {
    const int chunk = ring_chunk(interleaved, io);
    pf = interleaved ? 1 : 2;              // (measured: the frame loop is not fully unrolled, a deeper queue costs moves)
    Ir k = split_long_lines(ir, kRegLineDepth, (pf + 1) * chunk, ring);
    if ((int)ring.taps.size() * chunk * pf > 64) {
        pf = 1;
        k = split_long_lines(ir, kRegLineDepth, (pf + 1) * chunk, ring);
    }
    return k;
}

std::string jit_source(const Ir& ir, bool exact, bool interleaved, bool uniform, unsigned synth_mask, int io,
                       int n_ring_in = 0, int n_ring_out = 0, int ring_pf = 1, bool seg = false) {
    std::ostringstream src;
    src << "#define ZG_SYNTH_MASK " << synth_mask << "u\n";
    src << zg_stream_cuh_source << "\n";
    src << generate_tick_source(ir, exact, "ZgTick", n_ring_in, n_ring_out, ring_pf) << "\n";
    src << "extern \"C\" __global__ void __launch_bounds__(512, 1) zg_graph_kernel("
           "const __grid_constant__ zgk::StreamArgs a) {\n"
           "    zgk::stream_block<ZgTick, "
        << (interleaved ? "true" : "false") << ", " << (uniform ? "true" : "false") << ", " << io << ", "
        << (seg ? "true" : "false") /* built with time segments (kSeg) */ << ">(a);\n}\n";
    return src.str();
}

// source -> sm_100a cubin.  Needs libnvrtc only (no device, no driver).  Compiling takes ~0.5-1 s, so the cubins
// of the last few dozen distinct kernels are kept for the life of the process (key: the source text, which
// contains every specialisation: graph, layout, storage, parameter passing; plus the contraction flag) --
// a second plan of the same graph, e.g. one per rank-local stream or per test, costs a module load only.
struct CubinCache {
    std::mutex mu;
    std::vector<std::pair<std::string, std::vector<char>>> entries;     // most recent last
    static constexpr size_t kMax = 48;
};
CubinCache& cubin_cache() {
    static CubinCache c;
    return c;
}

int jit_cubin_uncached(const std::string& text, bool exact, std::vector<char>& cubin);

// Optional second level on disk: ZG_KERNEL_CACHE_DIR=<dir> keeps the cubins across processes (a service that
// restarts with the same graphs pays the NVRTC compile once).  File name = two independent 64-bit FNV-1a hashes of
// the key (contraction flag + NVRTC version + specialised source); the file repeats the key length and a third
// hash, so a truncated or foreign file is ignored, never loaded.  Written to a temporary name and renamed.
uint64_t fnv1a(const std::string& s, uint64_t h) {
    for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
    return h;
}
uint64_t fnv1a_bytes(const std::vector<char>& v) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (char c : v) { h ^= (unsigned char)c; h *= 0x100000001b3ull; }
    return h;
}
struct DiskKey {
    std::string path;
    uint64_t len, check;
};
bool disk_key(const std::string& key, DiskKey& k) {
    const char* dir = std::getenv("ZG_KERNEL_CACHE_DIR");
    if (!dir || !*dir) return false;
    char name[64];
    std::snprintf(name, sizeof name, "/zg_%016llx%016llx.cubin", (unsigned long long)fnv1a(key, 0xcbf29ce484222325ull),
                  (unsigned long long)fnv1a(key, 0x84222325cbf29ce4ull));
    k.path = std::string(dir) + name;
    k.len = key.size();
    k.check = fnv1a(key, 0x9e3779b97f4a7c15ull);
    return true;
}
bool disk_load(const DiskKey& k, std::vector<char>& cubin) {
    FILE* f = std::fopen(k.path.c_str(), "rb");
    if (!f) return false;
    uint64_t head[5] = {0, 0, 0, 0, 0};                       // magic, key length, check hash, cubin bytes, hash of the cubin
    bool ok = std::fread(head, sizeof head, 1, f) == 1 && head[0] == 0x32434755425a47ull /* file magic, format 2 */ &&
              head[1] == k.len && head[2] == k.check && head[3] > 0 && head[3] < (64u << 20);
    if (ok) {
        cubin.resize(head[3]);
        ok = std::fread(cubin.data(), 1, cubin.size(), f) == cubin.size() && std::fgetc(f) == EOF &&
             fnv1a_bytes(cubin) == head[4];                   // a damaged body is recompiled and rewritten, never loaded
    }
    std::fclose(f);
    if (!ok) cubin.clear();
    return ok;
}
void disk_store(const DiskKey& k, const std::vector<char>& cubin) {
    const std::string tmp = k.path + ".tmp" + std::to_string((long)getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;                                           // an unwritable cache directory is not an error
    const uint64_t head[5] = {0x32434755425a47ull, k.len, k.check, cubin.size(), fnv1a_bytes(cubin)};
    const bool ok = std::fwrite(head, sizeof head, 1, f) == 1 && std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), k.path.c_str()) != 0) std::remove(tmp.c_str());
}

int jit_cubin(const std::string& text, bool exact, std::vector<char>& cubin) {
    const std::string key = (exact ? "E" : "F") + text;
    CubinCache& c = cubin_cache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        for (auto& e : c.entries)
            if (e.first == key) { cubin = e.second; return ZG_OK; }
    }
    DiskKey dk;
    int major = 0, minor = 0;
    if (nvrtc().ok && nvrtc().version) nvrtc().version(&major, &minor);
    const bool on_disk = disk_key("nvrtc" + std::to_string(major) + "." + std::to_string(minor) + key, dk);
    if (!(on_disk && disk_load(dk, cubin))) {
        int st = jit_cubin_uncached(text, exact, cubin);
        if (st != ZG_OK) return st;
        if (on_disk) disk_store(dk, cubin);
    }
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.entries.size() >= CubinCache::kMax) c.entries.erase(c.entries.begin());
    c.entries.emplace_back(key, cubin);
    return ZG_OK;
}

int jit_cubin_uncached(const std::string& text, bool exact, std::vector<char>& cubin) {
    NvtxRange range("zg: NVRTC compile of a generated kernel");
    Nvrtc& n = nvrtc();
    if (!n.ok) return fail(ZG_ERR_CUDA, n.why);
    void* prog = nullptr;
    int r = n.createProgram(&prog, text.c_str(), "zg_graph_kernel.cu", 0, nullptr, nullptr);
    if (r != 0) return fail(ZG_ERR_CUDA, std::string("nvrtcCreateProgram: ") + n.getErrorString(r));
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--generate-line-info",
                          exact ? "--fmad=false" : "--fmad=true"};
    r = n.compileProgram(prog, 4, opts);
    if (r != 0) {
        size_t ls = 0;
        n.getProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.getProgramLog(prog, &log[0]);
        n.destroyProgram(&prog);
        return fail(ZG_ERR_CUDA, std::string("NVRTC could not compile the kernel for this graph: ") +
                                     n.getErrorString(r) + "\n" + log);
    }
    size_t cs = 0;
    n.getCUBINSize(prog, &cs);
    cubin.resize(cs);
    n.getCUBIN(prog, cubin.data());
    n.destroyProgram(&prog);
    return ZG_OK;
}

int jit_compile(zg_plan* p, bool uniform, Variant& v) {
    Driver& d = driver();
    if (!d.ok) return fail(ZG_ERR_CUDA, d.why);
    std::vector<char> cubin;
    int st = jit_cubin(jit_source(p->kir, p->exact, p->interleaved, uniform, p->synth_mask, p->io, (int)p->ring.taps.size(),
                                  (int)p->ring.out_lines.size(), p->ring_pf, p->seg_now),
                       p->exact, cubin);
    if (st != ZG_OK) return st;
    CUresult cr = d.moduleLoadData(&v.module, cubin.data());
    if (cr != CUDA_SUCCESS) return fail(ZG_ERR_CUDA, "cuModuleLoadData: " + cu_err(cr));
    cr = d.moduleGetFunction(&v.function, v.module, "zg_graph_kernel");
    if (cr != CUDA_SUCCESS) return fail(ZG_ERR_CUDA, "cuModuleGetFunction: " + cu_err(cr));
    d.funcGetAttribute(&v.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, v.function);
    v.ready = true;
    return ZG_OK;
}

int tune_env(const char* name) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : 0;
}

bool variant_is_sym(const zg_plan* p) {
    return p->sym_now && p->is_biquad && !p->opts.force_jit && p->lanes_now == 1 && p->exact && !tune_env("ZG_TUNE_NO_SYM");
}
int variant_index(const zg_plan* p) {
    const int lane_per_channel_of_k1b = p->lanes > 1 && p->lanes_now == 1 ? 4 : 0;     // never EXACT, hence never sym
    return (p->uniform_now ? 1 : 0) + (variant_is_sym(p) ? 2 : 0) + lane_per_channel_of_k1b + (p->seg_now ? 6 : 0);
}

int get_variant(zg_plan* p, bool uniform, Variant*& out) {
    const int vi = variant_index(p);
    Variant& v = p->variant[vi];
    out = &v;
    if (v.ready) return ZG_OK;
    if (p->is_fir) {
        cudaFuncAttributes fa;
        ZG_CUDA(cudaFuncGetAttributes(&fa, p->fir_tc ? (const void*)zg_fir_tc_kernel : fir_kernel_for(p->exact, p->interleaved)));
        v.regs = fa.numRegs;
        v.ready = true;
        return ZG_OK;
    }
    if (p->is_biquad && !p->opts.force_jit) {
        v.prebuilt = p->lanes_now > 1 ? biquad_lanes_kernel_for(p->bq.sections, p->exact, uniform)
                                      : biquad_kernel_for(p->bq.sections, p->exact, p->interleaved, uniform, variant_is_sym(p), p->seg_now);
        if (!v.prebuilt) return fail(ZG_ERR_INTERNAL, "no prebuilt biquad kernel for this section count");
        cudaFuncAttributes fa;
        ZG_CUDA(cudaFuncGetAttributes(&fa, (const void*)v.prebuilt));
        v.regs = fa.numRegs;
        v.ready = true;
        return ZG_OK;
    }
    return jit_compile(p, uniform, v);
}

// ---- parameters ------------------------------------------------------------------------------------

// value of kernel parameter slot j for channel c (c = -1: the scalar)
int sync_params(zg_plan* p) {
    if (!p->params_dirty) return ZG_OK;
    p->scan_dirty = true;                   // the tick's A matrix is made of these values
    p->AL_len = 0;
    if (p->is_fir) {
        // taps are shared by all channels: literals, or scalar $k parameters
        std::vector<float> taps(p->fir.taps.size());
        for (size_t k = 0; k < taps.size(); ++k) {
            const BiquadCoef& c = p->fir.taps[k];
            if (c.is_param && (p->h_params[c.param].size() != 1 || p->param_on_device[c.param]))
                return fail(ZG_ERR_UNSUPPORTED, "the FIR kernel shares its taps between channels: $" +
                                                    std::to_string(c.param) + " must be a scalar");
            taps[k] = c.is_param ? p->h_params[c.param][0] : c.value;
        }
        if (!p->d_taps) ZG_CUDA(cudaMalloc(&p->d_taps, taps.size() * sizeof(float)));
        ZG_CUDA(cudaMemcpy(p->d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
        p->uniform_now = true;
        p->params_dirty = false;
        return ZG_OK;
    }
    const int NP = p->kernel_n_param;
    bool uniform = NP <= zgk::kMaxUniform;
    for (auto& h : p->h_params) uniform = uniform && h.size() == 1;
    for (char d : p->param_on_device) uniform = uniform && !d;
    // kernel slot j -> (graph parameter | literal)
    auto slot_src = [&](int j, int& param, float& lit) {
        if (p->is_biquad && !p->opts.force_jit) {
            const BiquadCoef& c = p->bq.coef[j / 5][j % 5];
            param = c.is_param ? c.param : -1;
            lit = c.value;
        } else {
            param = j;
            lit = 0.f;
        }
    };
    p->sym_now = false;
    if (uniform) {
        for (int j = 0; j < NP; ++j) {
            int prm; float lit;
            slot_src(j, prm, lit);
            p->uparams[j] = prm >= 0 ? p->h_params[prm][0] : lit;
        }
        if (p->is_biquad && !p->opts.force_jit) {
            bool sym = true;
            for (int k = 0; k < p->bq.sections; ++k)
                sym = sym && std::memcmp(&p->uparams[5 * k], &p->uparams[5 * k + 2], sizeof(float)) == 0;
            p->sym_now = sym;
        }
    } else if (NP > 0) {
        if (p->is_biquad && !p->opts.force_jit) {
            // per-channel coefficients: b0 and b2 of a section are the same literal, or two parameters holding the
            // same values for every channel (every RBJ low-pass / high-pass / notch section, whatever f and Q)
            bool sym = true;
            for (int k = 0; k < p->bq.sections && sym; ++k) {
                const BiquadCoef& b0 = p->bq.coef[k][0];
                const BiquadCoef& b2 = p->bq.coef[k][2];
                if (b0.is_param != b2.is_param) sym = false;
                else if (!b0.is_param) sym = std::memcmp(&b0.value, &b2.value, sizeof(float)) == 0;
                else {
                    const std::vector<float>&u = p->h_params[b0.param], &w = p->h_params[b2.param];
                    sym = !p->param_on_device[b0.param] && !p->param_on_device[b2.param] && u.size() == w.size() &&
                          std::memcmp(u.data(), w.data(), u.size() * sizeof(float)) == 0;
                }
            }
            p->sym_now = sym;
        }
        std::vector<float> host((size_t)NP * p->ch_stride, 0.f);
        for (int j = 0; j < NP; ++j) {
            int prm; float lit;
            slot_src(j, prm, lit);
            float* row = host.data() + (size_t)j * p->ch_stride;
            if (prm < 0) std::fill(row, row + p->C, lit);
            else if (p->param_on_device[prm]) continue;                    // copied device to device below
            else if (p->h_params[prm].size() == 1) std::fill(row, row + p->C, p->h_params[prm][0]);
            else std::copy(p->h_params[prm].begin(), p->h_params[prm].end(), row);
        }
        if (!p->d_params) ZG_CUDA(cudaMalloc(&p->d_params, host.size() * sizeof(float)));
        ZG_CUDA(cudaMemcpy(p->d_params, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
        for (int j = 0; j < NP; ++j) {
            int prm; float lit;
            slot_src(j, prm, lit);
            if (prm >= 0 && p->param_on_device[prm])
                ZG_CUDA(cudaMemcpy(p->d_params + (size_t)j * p->ch_stride, p->d_user_params + (size_t)prm * p->ch_stride,
                                   p->C * sizeof(float), cudaMemcpyDeviceToDevice));
        }
    }
    p->uniform_now = uniform;
    p->params_dirty = false;
    return ZG_OK;
}

// ---- tensor maps ---------------------------------------------------------------------------------------

int encode_map(zg_plan* p, zgk::TensorMap* out, const void* base, int64_t C, int64_t T, int64_t ld, int box_rows) {
    Driver& d = driver();
    CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(out);
    // planar  [C][ld]: dim0 = sample (contiguous), dim1 = channel; rows of a box are swizzled (128B)
    // interleaved [T][ld] frames: dim0 = channel (contiguous), dim1 = sample
    cuuint64_t dims[2] = {(cuuint64_t)(p->interleaved ? C : T), (cuuint64_t)(p->interleaved ? T : C)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * p->io};
    // planar rows are 128 bytes of samples (32 fp32 / 64 bf16); interleaved rows are 32 channels
    cuuint32_t box[2] = {(cuuint32_t)(p->interleaved ? 32 : 128 / p->io), (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (int t = tune_env("ZG_TUNE_L2PROMO"))
        promo = t == 1 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : t == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : t == 4 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = d.tensorMapEncodeTiled(tm, p->io == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                        const_cast<void*>(base), dims, strides,
                                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        p->interleaved ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                                        promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZG_ERR_CUDA, "cuTensorMapEncodeTiled: " + cu_err(r));
    return ZG_OK;
}

// K1b whole-tile map: the planar [C][ld] buffer seen as {32 samples, C channels, T/32 boxes} (strides
// ld*4 and 128 bytes), box {32, rows, NB}: one TMA operation moves NB consecutive boxes of `rows`
// channels and lays them out in shared memory box after box, exactly like NB single-box operations.
// Only the full 32-sample boxes are covered (a ragged tail goes through the 2-D map, which clips at T).
bool encode_map_tile3d(zgk::TensorMap* out, const void* base, int64_t C, int64_t T, int64_t ld, int box_rows, int NB) {
    Driver& d = driver();
    if (T < 32) return false;
    cuuint64_t dims[3] = {32, (cuuint64_t)C, (cuuint64_t)(T / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)NB};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = d.tensorMapEncodeTiled(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                        const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;       // a driver that rejects the stride order: the kernel moves box by box
}

// ---- launch geometry ---------------------------------------------------------------------------------
//
// All warps of the launch should be resident at once (one wave) and spread evenly over the SMs:
// a CTA is `wpc` warps, every warp owns S stages of NT wires x NB boxes of 4 KB.  Shared memory, not
// registers, is what limits residency (the skeleton keeps 1 KB of slack to align the tiles).
// ZG_TUNE_WPC / ZG_TUNE_STAGES / ZG_TUNE_BOXES override the choice (tuning experiments only).
struct Geometry {
    int wpc, grid, stages, boxes, smem;
};

Geometry choose_geometry(const zg_plan* p, int64_t n_warps, int NT, int regs, int64_t T, bool segmented = false) {
    Geometry g{};
    const int BT = zgk::box_samples(p->interleaved, p->io), BB = zgk::box_bytes(p->interleaved, p->io);
    const int64_t per_sm = (n_warps + p->sm_count - 1) / p->sm_count;
    const int reg_warps = std::max(1, 65536 / (32 * std::max(regs, 32)));   // warps/SM the register file allows
    int wpc = (int)std::min<int64_t>({per_sm, 16, (int64_t)reg_warps});
    wpc = std::max(wpc, 1);
    // boxes per stage: 2 (256 contiguous bytes per channel row in flight together) when that still
    // leaves 3 stages, else 1
    int NB = 2;
    // A tick with little arithmetic is bound by HBM, and HBM likes fewer, longer streams: 7 resident warps
    // per SM with 4 boxes (512 contiguous bytes per channel row) per stage beat 14 x 2 by 3-11 % (copy graph
    // 5.72 -> 5.87 TB/s, osc >> LP 4.97 -> 5.51, FMA biquad x4 5.53 -> 5.74; profiles/r01_sweep_synth.jsonl).
    // Arithmetic-heavy ticks are issue-bound and keep all the warps the shared memory allows.
    // Planar blocks keep gaining from that trade up to ~4.25 instructions per byte -- the product-reusing EXACT
    // 4-section tick (32 instructions / 8 B): 0.791 ms against 0.805-0.845 with 14 x 2; at 8 sections (8 per byte)
    // or on interleaved frames 14 x 2 is the faster one (profiles/r01_sweep_sym2.jsonl).
    const int ops = p->tick_ops - (variant_is_sym(p) ? p->bq.sections : 0);
    const int bytes = p->io * std::max(1, p->n_buf_in + p->ir.n_out);
    // (ring reads of long delay lines are plain loads, hidden by other warps only: no trade of warps for run length)
    const bool long_runs = !p->interleaved && 4 * ops < 17 * bytes && !p->ring.any();
    if (long_runs && per_sm >= 7) {
        wpc = std::min(wpc, 7);
        NB = 4;
    }
    // More warps than one wave holds (one CTA per SM: shared memory is sized to fill it): the launch takes
    // waves x (time of one CTA ~ its warps, they share the SM's issue slots).  131 072 channels = 4096 warps: 16 per
    // CTA is 256 CTAs = 1.73 waves with a quarter of the SMs idle in the second; 14 per CTA is 293 CTAs = 1.98 waves.
    if (per_sm > wpc) {
        int best = wpc;
        int64_t best_cost = INT64_MAX;
        for (int w = wpc; w >= std::max(2, wpc / 2); --w) {
            const int64_t ctas = (n_warps + w - 1) / w;
            const int64_t cost = (ctas + p->sm_count - 1) / p->sm_count * w;
            if (cost < best_cost) { best_cost = cost; best = w; }
        }
        wpc = best;
    }
    if (int w = tune_env("ZG_TUNE_WPC")) wpc = std::min(std::max(w, 1), 16);
    const int budget = p->max_smem_optin - 1024 /*alignment slack*/ - 16 * 8 * 8 /*barriers*/;
    if (int b = tune_env("ZG_TUNE_BOXES")) NB = std::min(std::max(b, 1), 8);
    if (segmented) NB = NB >= 4 ? 4 : NB >= 2 ? 2 : 1;       // segment boundaries are multiples of 4 boxes
    NB = (int)std::min<int64_t>(NB, std::max<int64_t>(1, (T + BT - 1) / BT));
    auto stages_for = [&](int w, int nb) { return budget / (w * NT * nb * BB); };
    while (NB > 1 && stages_for(wpc, NB) < 2) --NB;
    int S = stages_for(wpc, NB);
    while (S < 2 && wpc > 1) {            // too many wires for that many warps: fewer warps per CTA
        --wpc;
        S = stages_for(wpc, NB);
    }
    S = std::min(S, 8);
    if (int st = tune_env("ZG_TUNE_STAGES")) S = std::min(std::max(st, 2), S);
    S = std::max(S, 2);
    g.wpc = wpc;
    g.stages = S;
    g.boxes = NB;
    g.grid = (int)((n_warps + wpc - 1) / wpc);
    g.smem = wpc * S * NT * NB * BB + 1024 + wpc * S * 8;
    return g;
}

// K1b: a box is (32 / lanes) channels x 32 samples; tiles are long in time (the pipeline fills and
// drains once per tile) and three stages deep.
Geometry choose_geometry_lanes(const zg_plan* p, int64_t n_warps, int regs, int64_t T) {
    Geometry g{};
    const int box_bytes = (32 / p->lanes) * 128;
    const int64_t per_sm = (n_warps + p->sm_count - 1) / p->sm_count;
    const int reg_warps = std::max(1, 65536 / (32 * std::max(regs, 32)));
    int wpc = (int)std::max<int64_t>(1, std::min<int64_t>({per_sm, 16, (int64_t)reg_warps}));
    if (int w = tune_env("ZG_TUNE_WPC")) wpc = std::min(std::max(w, 1), 16);
    const int budget = p->max_smem_optin - 1024 - 16 * 8 * 8;
    int S = 3;                                           // >= 3: a slot is refilled one tile after its store
    if (int st = tune_env("ZG_TUNE_STAGES")) S = std::min(std::max(st, 3), 8);
    int NB = 16;
    if (int b = tune_env("ZG_TUNE_BOXES")) NB = std::min(std::max(b, 1), 32);
    // The pipeline of the lanes runs through tile boundaries: the first lane of a channel is DRAIN + PD
    // iterations ahead of the last one, so a slot can only be refilled a few boxes into the next tile and
    // the ring (S stages of NB boxes) must be long enough for that: NB >= 8 whenever the block has more tiles
    // than stages (kernels/zg_biquad_lanes.cuh).  Fewer warps per CTA rather than shorter tiles.
    const int min_nb = 8;
    NB = std::max(NB, min_nb);
    NB = (int)std::min<int64_t>(NB, std::max<int64_t>(1, (T + zgk::kTileT - 1) / zgk::kTileT));
    auto need = [&](int w, int nb) { return w * (S * nb * box_bytes + zgk::kLanesXbufBytes); };
    while (NB > min_nb && need(wpc, NB) > budget) NB = std::max(NB / 2, min_nb);
    while (wpc > 1 && need(wpc, NB) > budget) --wpc;
    g.wpc = wpc;
    g.stages = S;
    g.boxes = NB;
    g.grid = (int)((n_warps + wpc - 1) / wpc);
    g.smem = need(wpc, NB) + 1024 + wpc * S * 8;
    return g;
}

// ---- time segments (few, long channels; FAST mode; linear ticks) -------------------------------------------------
//
// A lane-per-channel launch of C channels is C/32 warps however long the block is: BASELINE configs[1] (4096 channels
// x 65 536 samples) is 128 warps on 148 SMs.  A linear tick can be cut in time instead:
//   warm-up form  one launch, 8 + 4*K/L bytes per sample: segment g starts K samples early from zero state and
//                 discards those outputs.  Valid when every channel's |A^K|_inf <= 2^-30 (K = scan_warm, from the
//                 parameter values, float64): what is left of the true state after K ticks is below fp32 resolution.
//   two-pass form any linear / affine tick (poles on or outside the unit circle included), 12 bytes per sample:
//                 pass 1 from zero state -> boundary fix-up with A^L -> pass 2 from the true states.
// Both re-associate the arithmetic in time, so they are FAST-mode only; EXACT stays serial per channel.
struct Segments {
    int mode = 0;       // 0 none, 1 warm-up, 2 two-pass
    int n = 1, len = 0, warm = 0;
};
int segments_count(const Segments* sg) { return sg->n; }
int segments_warm(const Segments* sg) { return sg->warm; }

// per-channel parameter values on the host: row k has 1 (scalar) or C values
int host_params(zg_plan* p, std::vector<std::vector<float>>& rows) {
    rows = p->h_params;
    for (size_t k = 0; k < rows.size(); ++k) {
        if (!p->param_on_device[k]) continue;
        rows[k].resize(p->C);
        ZG_CUDA(cudaMemcpy(rows[k].data(), p->d_user_params + k * (size_t)p->ch_stride, p->C * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return ZG_OK;
}

constexpr double kScanTol = 9.3132257461547852e-10;     // 2^-30
constexpr int kScanMaxWarm = 8192;
constexpr int64_t kScanMaxChannels = 65536;             // beyond this a launch has enough warps without segments

// f(channel or -1, params of that channel) for the distinct parameter sets of the plan.  With per-channel parameters the
// channels are spread over the host's threads (20 us of float64 matrix work per channel: a parameter change on 16 384
// channels would otherwise stall the next launch for a third of a second); f must only touch per-channel results.
template <class F>
int for_each_param_set(zg_plan* p, F&& f) {
    std::vector<std::vector<float>> rows;
    int st = host_params(p, rows);
    if (st != ZG_OK) return st;
    bool per_channel = false;
    for (auto& r : rows) per_channel = per_channel || r.size() != 1;
    if (!per_channel) {
        std::vector<float> prm(rows.size());
        for (size_t k = 0; k < rows.size(); ++k) prm[k] = rows[k][0];
        f(-1, prm.data());
        return ZG_OK;
    }
    const int64_t C = p->C;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_threads = (int)std::min<int64_t>(std::min<unsigned>(hw, 32u), std::max<int64_t>(1, C / 256));
    auto work = [&](int64_t c0, int64_t c1) {
        std::vector<float> prm(rows.size());
        for (int64_t c = c0; c < c1; ++c) {
            for (size_t k = 0; k < rows.size(); ++k) prm[k] = rows[k].size() == 1 ? rows[k][0] : rows[k][c];
            f((int)c, prm.data());
        }
    };
    if (n_threads <= 1) {
        work(0, C);
        return ZG_OK;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, C * t / n_threads, C * (t + 1) / n_threads);
    for (auto& th : pool) th.join();
    return ZG_OK;
}

int scan_analyse(zg_plan* p) {
    p->scan_dirty = false;
    p->scan_warm = 0;
    if (!p->scan_ok || p->C > kScanMaxChannels) return ZG_OK;
    const int unit = 4 * zgk::box_samples(p->interleaved, p->io);
    std::atomic<int> worst{unit};
    std::atomic<bool> all{true};
    int st = for_each_param_set(p, [&](int, const float* prm) {
        if (!all.load(std::memory_order_relaxed)) return;            // some channel never forgets: the answer is known
        std::vector<double> A;
        const int K = tick_matrix(p->ir, prm, A) ? decay_length(A, p->ir.n_state, unit, kScanMaxWarm, kScanTol) : 0;
        if (K == 0) all.store(false, std::memory_order_relaxed);
        int w = worst.load(std::memory_order_relaxed);
        while (K > w && !worst.compare_exchange_weak(w, K, std::memory_order_relaxed)) {}
    });
    if (st != ZG_OK) return st;
    p->scan_warm = all.load() ? worst.load() : 0;
    return ZG_OK;
}

// A^L on the device for the fix-up kernel
int scan_prepare_AL(zg_plan* p, int L) {
    if (p->AL_len == L) return ZG_OK;
    const int n = p->ir.n_state;
    bool per_channel = false;
    for (size_t k = 0; k < p->h_params.size(); ++k) per_channel = per_channel || p->h_params[k].size() != 1 || p->param_on_device[k];
    const size_t need = (size_t)n * n * (per_channel ? (size_t)p->ch_stride : 1);
    std::vector<float> host(std::max<size_t>(need, 1), 0.f);
    std::atomic<bool> finite{true};
    int st = for_each_param_set(p, [&](int c, const float* prm) {
        std::vector<double> A, AL;
        if (!tick_matrix(p->ir, prm, A)) { finite.store(false); return; }
        mat_pow(A, n, L, AL);
        for (int e = 0; e < n * n; ++e) {
            if (!std::isfinite(AL[e]) || std::fabs(AL[e]) > 3e38) { finite.store(false); return; }
            if (c < 0) host[e] = (float)AL[e];
            else host[(size_t)e * p->ch_stride + c] = (float)AL[e];
        }
    });
    if (st != ZG_OK) return st;
    if (!finite.load()) return fail(ZG_ERR_UNSUPPORTED, "the tick's state matrix overflows fp32 over one time segment (unstable graph): "
                                                  "use time_parallel = ZG_TP_OFF");
    if (need > p->AL_floats) {
        if (p->d_AL) cudaFree(p->d_AL);
        p->d_AL = nullptr;
        p->AL_floats = 0;
        ZG_CUDA(cudaMalloc(&p->d_AL, std::max<size_t>(need, 1) * sizeof(float)));
        p->AL_floats = need;
    }
    if (need) ZG_CUDA(cudaMemcpy(p->d_AL, host.data(), need * sizeof(float), cudaMemcpyHostToDevice));
    p->AL_len = L;
    p->AL_per_channel = per_channel;
    return ZG_OK;
}

bool tick_is_light(const zg_plan* p) {       // the HBM-bound side of choose_geometry's trade (7 warps x 4 boxes)
    const int bytes = p->io * std::max(1, p->n_buf_in + p->ir.n_out);
    return !p->interleaved && 4 * p->tick_ops < 17 * bytes;
}

int choose_segments(zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t c_count, Segments& sg) {
    sg = Segments{};
    const int tp = tune_env("ZG_TUNE_TP") ? tune_env("ZG_TUNE_TP") - 1 : p->opts.time_parallel;
    if (tp == ZG_TP_OFF || !p->scan_ok) return ZG_OK;
    const int unit = 4 * zgk::box_samples(p->interleaved, p->io);
    const int64_t groups = (c_count + 31) / 32;
    const int64_t target = (int64_t)(tick_is_light(p) ? 7 : 14) * p->sm_count;
    int64_t G = target / groups;
    if (int g = tune_env("ZG_TUNE_SEGS")) G = g;
    else if (tp == ZG_TP_AUTO && groups * 2 > target) return ZG_OK;         // enough channels for one lane each
    G = std::max<int64_t>(G, 2);
    if (p->scan_dirty) {
        int st = scan_analyse(p);
        if (st != ZG_OK) return st;
    }
    bool in_place = false;
    for (int o = 0; o < p->ir.n_out; ++o)
        for (int k = 0; k < p->ir.n_in; ++k)
            in_place = in_place || (!(p->synth_mask & (1u << k)) && out[o] == in[k]);
    int K = p->scan_warm;
    if (int w = tune_env("ZG_TUNE_WARM")) K = (w + unit - 1) / unit * unit;
    // the warm-up of a segment re-reads the end of the previous one, which an in-place block has overwritten by then
    const bool warm_form = K > 0 && !in_place && tp != ZG_TP_TWO_PASS;
    if (tp == ZG_TP_WARMUP && !warm_form)
        return fail(in_place ? ZG_ERR_ARG : ZG_ERR_UNSUPPORTED,
                    in_place ? "the warm-up form of time_parallel does not run in place (segments re-read their predecessor's samples)"
                             : "time_parallel = ZG_TP_WARMUP: with these parameter values some channel does not forget its state "
                               "within " + std::to_string(kScanMaxWarm) + " samples (use ZG_TP_TWO_PASS)");
    // K1b (sections across lanes) is as fast as the two-pass form, which moves 12 instead of 8 bytes per sample
    if (!warm_form && tp == ZG_TP_AUTO && p->lanes > 1) return ZG_OK;
    int64_t L = ((T + G - 1) / G + unit - 1) / unit * unit;
    if (warm_form) L = std::max<int64_t>(L, tp == ZG_TP_AUTO ? 8 * (int64_t)K : 2 * (int64_t)K);   // auto: <= 12.5 % more samples evaluated
    else L = std::max<int64_t>(L, 2 * unit);
    G = (T + L - 1) / L;
    if (G < 2) return ZG_OK;                                               // the block is too short to cut
    sg.mode = warm_form ? 1 : 2;
    sg.n = (int)G;
    sg.len = (int)L;
    sg.warm = warm_form ? K : 0;
    return ZG_OK;
}

// One kernel launch over channels [c_begin, c_begin + c_count) of the plan; in/out point at the first
// of those channels.  `advance` = this launch ends the block (the stream position moves on by T).
int launch_fir(zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t ld_in, int64_t ld_out,
               cudaStream_t stream, int64_t c_begin, int64_t c_count, bool advance);

// K1s (kernels/zg_biquad_split.cuh): biquad cascades with many channels, planar fp32, whole boxes.  Persistent CTAs of G
// groups of `sections / spw` warps; a group owns a ring of S tiles of NB boxes of 32 channel rows and works through a
// contiguous range of the tile sequence (row after row), every group the same number of tiles: 65 536 channels x 8192
// samples are 2048 rows over 3 x 148 groups = 4.61 rows each, so a range begins and ends in the middle of a row and the
// delay lines of that row travel from one group to the next through HBM.  Ranges are at least two rows long (fewer
// groups otherwise): a row is cut at most once.  Returns false when the launch should stay on K1.
struct SplitGeometry {
    int spw, wpg, groups, stages, boxes, grid, smem, hand_boxes;
    int n_segs, seg_boxes, warm_boxes;      // time segments (FAST, warm-up form): rows are (channel group, segment)
};

struct Segments;
bool choose_split(const zg_plan* p, int64_t T, int64_t c_count, SplitGeometry& g, const Segments* sg = nullptr);
int segments_count(const Segments* sg);
int segments_warm(const Segments* sg);

bool choose_split(const zg_plan* p, int64_t T, int64_t c_count, SplitGeometry& g, const Segments* sg) {
    // zg_plan_opts.section_warps: 0 = auto, 1 = never, 2 = whenever the shape allows (ZG_TUNE_SPLIT overrides: 1 / 2 / 3)
    const int mode = tune_env("ZG_TUNE_SPLIT") ? tune_env("ZG_TUNE_SPLIT") - 1 : p->opts.section_warps;
    if (mode == 1 || !p->is_biquad || p->opts.force_jit || p->interleaved || p->io != 4) return false;
    if (T % zgk::kTileT != 0 || T < 4 * zgk::kTileT || T > (1ll << 30)) return false;
    // a plan sized for K1b (few channels): only when the auto rule, not the caller, chose the lanes
    if (p->lanes > 1 && (p->opts.lanes_per_channel != 0 || tune_env("ZG_TUNE_LANES"))) return false;
    const int S = p->bq.sections;
    // with time segments a row of the tile sequence is (channel group, segment)
    // (as few segments as give three groups per CTA a row each: every segment pays its warm-up; BASELINE configs[1]:
    //  4 segments 0.366 ms, 8 segments 0.375)
    const int want_segs = sg ? (int)std::min<int64_t>(segments_count(sg),
                                                      std::max<int64_t>(2, (3 * (int64_t)p->sm_count + (c_count + 31) / 32 - 1) / ((c_count + 31) / 32)))
                             : 1;
    int64_t n_cg = (c_count + 31) / 32 * want_segs;
    int64_t row_boxes = T / zgk::kTileT;
    if (sg) row_boxes = std::max<int64_t>(8, (row_boxes - segments_warm(sg) / zgk::kTileT) / want_segs + segments_warm(sg) / zgk::kTileT);
    g.n_segs = 1;
    g.seg_boxes = g.warm_boxes = 0;
    // sections per warp and groups per CTA, measured on 65 536 x 8192 (EXACT, GB/s of 8 B/sample; K1 in brackets):
    //   3 sections: 1 x 3 groups 6306 (5345)   4: 1 x 3 6087 (5555)   5: 1 x 3 5078 (4780)   6: 2 x 4 4893 (4096)
    //   7: 1 x 2 3486 (3281)   8: 2 x 4 3844 (2748);   2 sections: 5791 (5977) -- two warps per group have nothing to
    //   overlap, K1 stays.  Three groups of single-section warps put three independent recurrences on every scheduler
    //   (EXACT: 3 dependent instructions of the 8 per sample and section).  FAST with 3 sections: no gain, K1 stays.
    int spw = S == 6 || S == 8 ? 2 : 1;
    int groups = S == 6 || S == 8 ? 4 : S == 7 ? 2 : 3;
    if (int w = tune_env("ZG_TUNE_SPLIT_SPW")) spw = w;
    if (spw < 1 || S % spw != 0 || !biquad_split_kernel_for(S, spw, p->exact, variant_is_sym(p), p->uniform_now)) return false;
    if (mode == 0 && (S < 3 || (S == 3 && !p->exact))) return false;
    g.spw = spw;
    g.wpg = S / spw;
    g.hand_boxes = 1;
    groups = std::max(1, std::min(groups, 16 / g.wpg));
    // every group's range is at least one row long, so that a row is cut at most once
    auto grid_for = [&](int G) { return (int)std::max<int64_t>(1, std::min<int64_t>(p->sm_count, n_cg / G)); };
    bool fits_k1s = true;                              // auto: is this the kernel for the shape?
    if (S == 4 && spw == 1) {
        // Fewer channel groups than 3 x SMs: fewer groups per CTA keep every SM busy.  A group advances one sample of its 32
        // channels in c(G) cycles when G groups share an SM -- measured (EXACT / FAST): c(1) = 19.9 / 17.7 (BASELINE
        // configs[1], four boxes per hand-over: every warp is alone on its scheduler and only its recurrence bounds it;
        // K1b: 21.3), c(2) = 29.5 / 25.6 (8192 x 32 768: 0.49 ms against K1b's 0.58), c(3) = 37.5 / 36 (HBM-bound from
        // there on: 16 384 x 16 384 runs at 0.91 of the copy peak, K1: 0.55) -- and a launch takes rows per group x c(G).
        const double cyc[2][3] = {{17.7, 25.6, 36.0}, {19.9, 29.5, 37.5}};
        double best = 0;
        int best_g = 0;
        for (int G = 1; G <= 3; ++G) {
            if (G > n_cg) break;
            const double cost = (double)n_cg / ((double)G * grid_for(G)) * cyc[p->exact ? 1 : 0][G - 1];
            if (!best_g || cost < best) { best = cost; best_g = G; }
        }
        groups = best_g;
        // below ~2.5 warps of channels per SM a FAST plan has K1b, which this kernel does not beat (uncut; cut in time
        // it is this kernel again, with (channel group, segment) rows)
        if (!p->exact && p->lanes > 1 && !sg) fits_k1s = false;
    } else if (p->lanes > 1 || sg) {
        fits_k1s = false;
    } else {
        // other section counts: as many groups per CTA (up to the measured optimum) as still give every SM its groups,
        // at least two -- one group per SM needs the several-boxes-per-hand-over form, which exists for 4 sections only.
        // Measured on 16 384 x 16 384 (EXACT, ms, K1 in brackets): 3 sections 0.356 (0.514), 5: 0.441 (0.802),
        // 6: 0.593 (0.890), 8: 0.606 (1.792); 9600 x 16 384: 3 sections 0.264 (0.509), 8: 0.420 (1.758)
        while (groups > 2 && n_cg < (int64_t)groups * p->sm_count) --groups;
        if (n_cg < (int64_t)groups * p->sm_count) fits_k1s = false;
    }
    if (int t = tune_env("ZG_TUNE_SPLIT_G")) groups = std::min(std::max(t, 1), 16 / g.wpg);
    g.groups = (int)std::max<int64_t>(1, std::min<int64_t>(groups, n_cg));
    g.grid = grid_for(g.groups);
    if (g.groups == 1 && S == 4 && spw == 1) {
        const int hb = tune_env("ZG_TUNE_SPLIT_HB") ? tune_env("ZG_TUNE_SPLIT_HB") : 4;
        if (hb > 1 && biquad_split_kernel_for(S, 1, p->exact, false, p->uniform_now, hb)) g.hand_boxes = hb;
    }
    // the ring: one group per SM runs its four warps four boxes apart (3 x 16 boxes); else two stages, as long as they fit
    g.stages = g.hand_boxes > 1 ? 3 : 2;
    if (int t = tune_env("ZG_TUNE_STAGES")) g.stages = std::min(std::max(t, 2), zgk::kSplitAckRing - 2);
    const int budget = p->max_smem_optin - 1024 /*alignment slack*/;
    int nb = g.hand_boxes > 1 ? 16 : 14;
    if (int t = tune_env("ZG_TUNE_BOXES")) nb = std::min(std::max(t, 1), 32);
    nb = (int)std::min<int64_t>(nb, row_boxes);
    auto need = [&](int boxes) {
        return g.groups * (g.stages * boxes * zgk::kTileBytes + zgk::split_group_extra_bytes(g.stages, boxes, g.wpg)) + 16;
    };
    while (nb > 1 && need(nb) > budget) --nb;
    if (need(nb) > budget) return false;
    if (g.hand_boxes > 1) {
        nb = nb / g.hand_boxes * g.hand_boxes;
        if (nb < 2 * g.hand_boxes) return false;
    } else if (sg) {
        if (nb >= 8 && !tune_env("ZG_TUNE_BOXES")) nb = 8;      // segment length and warm-up are rounded to whole tiles below
    } else if (!tune_env("ZG_TUNE_BOXES")) {
        // tiles that divide the row leave no ragged tile at its end (8192 samples: 8 boxes rather than 9)
        for (int d = nb; d >= std::max(2, nb - 2); --d)
            if (row_boxes % d == 0) { nb = d; break; }
    }
    g.boxes = nb;
    g.smem = need(nb) + 1024;
    if (sg) {
        // segment length and warm-up in whole tiles; the last segment takes the ragged end of the block
        const int64_t block_boxes = T / zgk::kTileT;
        const int64_t Kb = (segments_warm(sg) / zgk::kTileT + nb - 1) / nb * nb;
        if (block_boxes <= Kb + nb) return false;
        const int64_t Lb = std::max<int64_t>(2 * Kb, ((block_boxes - Kb + want_segs - 1) / want_segs + nb - 1) / nb * nb);
        const int64_t n = (block_boxes - Kb + Lb - 1) / Lb;
        if (n < 2 || !fits_k1s || S != 4 || p->ir.n_state != p->kernel_n_state) return false;
        g.n_segs = (int)n;
        g.seg_boxes = (int)Lb;
        g.warm_boxes = (int)Kb;
        n_cg = (c_count + 31) / 32 * n;
        row_boxes = Lb + Kb;
        g.groups = (int)std::max<int64_t>(1, std::min<int64_t>(g.groups, n_cg));
        g.grid = grid_for(g.groups);
        return true;
    }
    if (mode == 2) return true;
    // auto: rows of at least eight tiles (the ring fills once per launch), and -- other than for 4 sections, where the
    // group count follows the channel count -- a GPU's worth of groups
    if ((row_boxes + nb - 1) / nb < 8) return false;
    return fits_k1s;
}

int launch_split(zg_plan* p, const SplitGeometry& g, const void* const* in, void* const* out, int64_t T, int64_t ld_in,
                 int64_t ld_out, cudaStream_t stream, int64_t c_begin, int64_t c_count, bool advance) {
    const bool sym = g.hand_boxes == 1 && variant_is_sym(p);
    SplitKernelPtr fn = biquad_split_kernel_for(p->bq.sections, g.spw, p->exact, sym, p->uniform_now, g.hand_boxes);
    zgk::SplitArgs a;
    std::memset(&a, 0, sizeof a);
    if (!encode_map_tile3d(&a.in_map, in[0], c_count, T, ld_in, 32, g.boxes) ||
        !encode_map_tile3d(&a.out_map, out[0], c_count, T, ld_out, 32, g.boxes))
        return fail(ZG_ERR_CUDA, "cuTensorMapEncodeTiled (whole-tile map of the section-split biquad kernel)");
    const size_t n_cg_plan = (size_t)(p->ch_stride + 31) / 32;
    if (!p->d_split_ctl || p->split_wpg_alloc < g.wpg || p->split_segs_alloc < g.n_segs) {
        if (p->d_split_ctl) cudaFree(p->d_split_ctl);
        if (p->d_split_flags) cudaFree(p->d_split_flags);
        if (p->d_split_carry) cudaFree(p->d_split_carry);
        p->d_split_ctl = nullptr;
        p->d_split_flags = nullptr;
        p->d_split_carry = nullptr;
        ZG_CUDA(cudaMalloc(&p->d_split_ctl, 4 * sizeof(unsigned long long)));
        ZG_CUDA(cudaMemset(p->d_split_ctl, 0, 4 * sizeof(unsigned long long)));
        const int wpg_alloc = std::max(p->split_wpg_alloc, g.wpg), segs_alloc = std::max(p->split_segs_alloc, g.n_segs);
        ZG_CUDA(cudaMalloc(&p->d_split_flags, n_cg_plan * wpg_alloc * segs_alloc * sizeof(unsigned long long)));
        ZG_CUDA(cudaMemset(p->d_split_flags, 0, n_cg_plan * wpg_alloc * segs_alloc * sizeof(unsigned long long)));
        ZG_CUDA(cudaMalloc(&p->d_split_carry, (size_t)p->kernel_n_state * 2 * p->ch_stride * segs_alloc * sizeof(float)));
        p->split_wpg_alloc = wpg_alloc;
        p->split_segs_alloc = segs_alloc;
    }
    a.state = p->d_state + c_begin;
    a.params = p->d_params ? p->d_params + c_begin : nullptr;
    a.ch_stride = p->ch_stride;
    a.channels = (int)c_count;
    a.n_samples = (int)T;
    a.stages = g.stages;
    a.boxes = g.boxes;
    a.ctl = p->d_split_ctl;
    a.flags = p->d_split_flags + (c_begin / 32) * g.wpg * g.n_segs;
    a.carry = p->d_split_carry + c_begin * g.n_segs;
    a.carry_stride = p->ch_stride * g.n_segs;
    a.state_out = a.state;
    if (g.n_segs > 1) {
        // the last segments leave the block's final state in another buffer (a first segment may still have to read the old
        // one), copied back behind the kernel
        if (!p->d_state_alt) ZG_CUDA(cudaMalloc(&p->d_state_alt, (size_t)std::max(p->ir.n_state, 1) * p->ch_stride * sizeof(float)));
        a.state_out = p->d_state_alt + c_begin;
        a.n_segs = g.n_segs;
        a.seg_boxes = g.seg_boxes;
        a.warm_boxes = g.warm_boxes;
    }
    for (int j = 0; j < p->kernel_n_state; ++j) a.state_row[j] = p->state_row[j];
    if (p->uniform_now) std::memcpy(a.uparams, p->uparams, sizeof(float) * std::min(p->kernel_n_param, zgk::kMaxUniform));
    int st = raise_max_smem((const void*)fn, p->opts.device, g.smem);
    if (st != ZG_OK) return st;
    if (p->split_fn != (const void*)fn) {
        cudaFuncAttributes fa;
        ZG_CUDA(cudaFuncGetAttributes(&fa, (const void*)fn));
        p->split_regs = fa.numRegs;
        p->split_fn = (const void*)fn;
    }
    p->split_spw = g.spw;
    void* args[] = {&a};
    ZG_CUDA(cudaLaunchKernel((const void*)fn, dim3(g.grid), dim3(g.groups * g.wpg * 32), args, g.smem, stream));
    if (g.n_segs > 1)
        ZG_CUDA(cudaMemcpy2DAsync(p->d_state + c_begin, p->ch_stride * sizeof(float), p->d_state_alt + c_begin,
                                  p->ch_stride * sizeof(float), (size_t)c_count * sizeof(float), (size_t)p->kernel_n_state,
                                  cudaMemcpyDeviceToDevice, stream));
    p->launches += 1;
    p->split_now = true;
    p->split_hand_boxes = g.hand_boxes;
    p->last_segs = g.n_segs;
    p->last_seg_mode = g.n_segs > 1 ? 1 : 0;
    p->last_seg_len = g.seg_boxes * zgk::kTileT;
    p->last_seg_warm = g.warm_boxes * zgk::kTileT;
    if (advance) p->stream_pos += T;
    p->last_smem = g.smem;
    p->last_threads = g.groups * g.wpg * 32;
    p->last_stages = g.stages;
    p->last_boxes = g.boxes;
    p->last_grid = g.grid;
    return ZG_OK;
}

int launch(zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t ld_in, int64_t ld_out,
           cudaStream_t stream, int64_t c_begin, int64_t c_count, bool advance) {
    Driver& d = driver();
    if (!d.ok) return fail(ZG_ERR_CUDA, d.why);
    int st = sync_params(p);
    if (st != ZG_OK) return st;
    if (p->is_fir) return launch_fir(p, in, out, T, ld_in, ld_out, stream, c_begin, c_count, advance);
    Segments sg;
    st = choose_segments(p, in, out, T, c_count, sg);
    if (st != ZG_OK) return st;
    p->split_now = false;
    if (sg.mode <= 1) {
        p->lanes_now = sg.mode ? 1 : p->lanes;
        p->seg_now = false;                 // (K1s is one kernel with and without segments)
        SplitGeometry sp{};
        if (choose_split(p, T, c_count, sp, sg.mode == 1 ? &sg : nullptr)) {
            p->lanes_now = 1;
            return launch_split(p, sp, in, out, T, ld_in, ld_out, stream, c_begin, c_count, advance);
        }
    }
    p->lanes_now = sg.mode ? 1 : p->lanes;
    p->seg_now = sg.mode != 0;
    Variant* v = nullptr;
    st = get_variant(p, p->uniform_now, v);
    if (st != ZG_OK) return st;

    zgk::StreamArgs a;
    std::memset(&a, 0, sizeof a);
    for (int k = 0; k < p->ir.n_in; ++k) {
        if (p->synth_mask & (1u << k)) continue;
        st = encode_map(p, &a.in_map[k], in[k], c_count, T, ld_in, 32 / p->lanes_now);
        if (st != ZG_OK) return st;
    }
    for (int o = 0; o < p->ir.n_out; ++o) {
        st = encode_map(p, &a.out_map[o], out[o], c_count, T, ld_out, 32 / p->lanes_now);
        if (st != ZG_OK) return st;
    }
    a.state = p->d_state + c_begin;
    a.params = p->d_params ? p->d_params + c_begin : nullptr;
    a.ch_stride = p->ch_stride;
    a.stream_pos = p->stream_pos;
    a.channels = (int)c_count;
    a.n_samples = (int)T;
    a.dirac_mask = p->dirac_mask;
    if (!p->is_biquad || p->opts.force_jit) {
        // generated kernel: short lines at fixed rows; window slots of long lines at ring rows that move with the
        // stream position (the value pushed `ago` ticks before this block starts), loaded but never written back
        for (int j = 0; j < p->kernel_n_state; ++j) {
            const KernelSlot& ks = p->ring.slots[j];
            if (ks.ring_depth == 0) { a.state_row[j] = ks.row; continue; }
            int64_t r = (p->stream_pos - ks.ago) % ks.ring_depth;
            if (r < 0) r += ks.ring_depth;
            a.state_row[j] = ks.row + (int)r;
            a.state_nowrite |= 1ull << j;
        }
        for (size_t r = 0; r < p->ring.taps.size(); ++r) {
            const IrLine& l = p->ir.lines[p->ring.taps[r].line];
            a.ring_in_row0[r] = l.offset;
            a.ring_in_depth[r] = l.depth;
            a.ring_in_delay[r] = p->ring.taps[r].n;
        }
        for (size_t w = 0; w < p->ring.out_lines.size(); ++w) {
            const IrLine& l = p->ir.lines[p->ring.out_lines[w]];
            a.ring_out_row0[w] = l.offset;
            a.ring_out_depth[w] = l.depth;
        }
    } else {
        for (int j = 0; j < p->kernel_n_state; ++j) a.state_row[j] = p->state_row[j];
    }
    if (p->uniform_now) std::memcpy(a.uparams, p->uparams, sizeof(float) * std::min(p->kernel_n_param, zgk::kMaxUniform));

    // wires per pipeline stage as the kernel lays them out: wire k of a stage belongs to input k / output k
    // (kernels/zg_stream.cuh: NT = max(N_IN, N_OUT)), whether or not input k is synthesised
    const int NT = std::max(1, std::max(p->ir.n_in, p->ir.n_out));
    const int cpw = 32 / p->lanes_now;                 // channels per warp
    const int64_t n_warps = (c_count + cpw - 1) / cpw * sg.n;               // (channel group, time segment)
    if (n_warps > 0x7fffffffLL / 32) return fail(ZG_ERR_ARG, "block too large for one launch");
    Geometry g = p->lanes_now > 1 ? choose_geometry_lanes(p, n_warps, v->regs, T)
                                  : choose_geometry(p, n_warps, NT, v->regs, sg.mode ? sg.len + sg.warm : T, sg.mode != 0);
    a.stages = g.stages;
    a.boxes = g.boxes;
    if (sg.mode) {
        a.n_segs = sg.n;
        a.seg_len = sg.len;
        a.seg_warm = sg.warm;
    }
    const int n_rows = std::max(p->ir.n_state, 1);
    if (sg.mode == 2) {
        const size_t need = (size_t)(sg.n - 1) * n_rows * p->ch_stride;
        if (need > p->seg_state_floats) {
            if (p->d_seg_state) cudaFree(p->d_seg_state);
            p->d_seg_state = nullptr;
            p->seg_state_floats = 0;
            ZG_CUDA(cudaMalloc(&p->d_seg_state, need * sizeof(float)));
            ZG_CUDA(cudaMemset(p->d_seg_state, 0, need * sizeof(float)));      // rows no kernel slot maps to stay zero
            p->seg_state_floats = need;
        }
        st = scan_prepare_AL(p, sg.len);
        if (st != ZG_OK) return st;
        a.seg_state = p->d_seg_state + c_begin;
        a.seg_state_stride = (long long)n_rows * p->ch_stride;
    }
    if (p->lanes_now > 1 && !tune_env("ZG_TUNE_NO3D")) {
        const bool ok = encode_map_tile3d(&a.in_map[1], in[0], c_count, T, ld_in, cpw, g.boxes) &&
                        encode_map_tile3d(&a.out_map[1], out[0], c_count, T, ld_out, cpw, g.boxes);
        a.flags = ok ? 1 : 0;
    }
    // K4 refill policy (measured, profiles/r01_sweep_refill.jsonl): a tick with a lot of arithmetic per sample
    // is bound by instruction issue and gains from refilling a slot in the middle of the next tile (the load
    // latency hides behind arithmetic); a light tick is bound by HBM and does better when a warp's load and
    // store requests leave together.  ZG_TUNE_LATE_REFILL=1 / =2 force late / early.
    bool late = p->light_tick;
    if (int t = tune_env("ZG_TUNE_LATE_REFILL")) late = t == 1;
    if (late) a.flags |= 2;
    if (int h = tune_env("ZG_TUNE_L2HINT")) a.flags |= (h & 3) << 2;       // 1 = loads, 2 = stores, 3 = both: evict-first
    // L2 prefetch of the input rows in long runs ahead of the 128-byte-wide TMA boxes (experiment, off by default):
    // ZG_TUNE_PF = window in bytes per channel row, ZG_TUNE_PFD = tiles ahead (default 2)
    if (int w = tune_env("ZG_TUNE_PF"); w > 0 && !p->interleaved && p->lanes_now == 1) {
        const int tile_bytes = g.boxes * 128;
        a.pf_window = std::max(tile_bytes, w / tile_bytes * tile_bytes);
        a.pf_dist = tune_env("ZG_TUNE_PFD") > 0 ? tune_env("ZG_TUNE_PFD") : 2;
        a.in_pitch_bytes = ld_in * p->io;
        for (int k = 0; k < p->ir.n_in; ++k) a.in_base[k] = static_cast<const unsigned char*>(in[k]);
        a.flags |= 16;
    }

    if (v->prebuilt) {
        st = raise_max_smem((const void*)v->prebuilt, p->opts.device, g.smem);
        if (st != ZG_OK) return st;
    } else if (g.smem > v->max_smem_set) {                  // a generated kernel is this plan's own module
        CUresult cr = d.funcSetAttribute(v->function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, g.smem);
        if (cr != CUDA_SUCCESS) return fail(ZG_ERR_CUDA, "cuFuncSetAttribute(max dynamic smem): " + cu_err(cr));
        v->max_smem_set = g.smem;
    }
    auto launch_once = [&](int pass_flag) -> int {
        zgk::StreamArgs b = a;
        b.flags |= pass_flag;
        void* args[] = {&b};
        if (v->prebuilt) {
            ZG_CUDA(cudaLaunchKernel((const void*)v->prebuilt, dim3(g.grid), dim3(g.wpc * 32), args, g.smem, stream));
        } else {
            CUresult cr = d.launchKernel(v->function, g.grid, 1, 1, g.wpc * 32, 1, 1, g.smem, (CUstream)stream, args, nullptr);
            if (cr != CUDA_SUCCESS) return fail(ZG_ERR_CUDA, "cuLaunchKernel: " + cu_err(cr));
        }
        p->launches += 1;
        return ZG_OK;
    };
    if (sg.mode == 2) {
        // pass 1: final states from zero; fix-up: true boundary states; pass 2: the samples
        st = launch_once(32);
        if (st != ZG_OK) return st;
        if (sg.n > 2) {
            const int threads = 128;
            zg_scan_fixup_kernel<<<(unsigned)((c_count + threads - 1) / threads), threads, 0, stream>>>(
                a.seg_state, a.seg_state_stride, p->ch_stride, (int)c_count, p->ir.n_state, sg.n - 1,
                p->AL_per_channel ? p->d_AL + c_begin : p->d_AL, p->AL_per_channel ? 1 : 0);
            ZG_CUDA(cudaGetLastError());
            p->launches += 1;
        }
        st = launch_once(64);
    } else {
        st = launch_once(0);
    }
    if (st != ZG_OK) return st;
    p->last_segs = sg.n;
    p->last_seg_mode = sg.mode;
    p->last_seg_len = sg.len;
    p->last_seg_warm = sg.warm;
    if (advance) p->stream_pos += T;
    p->last_smem = g.smem;
    p->last_threads = g.wpc * 32;
    p->last_stages = g.stages;
    p->last_boxes = g.boxes;
    return ZG_OK;
}

// K3 geometry: W warps per CTA (one 32-sample box each per step), ring of H + 2W input boxes, 2 output
// boxes per warp; time is cut into segments when there are fewer channel groups than ~2 CTAs per SM.
// K3t: persistent CTAs, one per SM, each a contiguous range of (128-channel group, 128-sample output tile) items
int launch_fir_tc(zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t ld_in, int64_t ld_out,
                  cudaStream_t stream, int64_t c_begin, int64_t c_count, bool advance) {
    const int N = (int)p->fir.taps.size();
    zgk::FirTcArgs a;
    std::memset(&a, 0, sizeof a);
    int st = encode_map(p, &a.in_map, in[0], c_count, T, ld_in, 128);
    if (st != ZG_OK) return st;
    st = encode_map(p, &a.out_map, out[0], c_count, T, ld_out, 128);
    if (st != ZG_OK) return st;
    a.state_in = p->d_state + c_begin;
    a.taps = p->d_taps;
    a.ch_stride = p->ch_stride;
    a.channels = (int)c_count;
    a.n_samples = (int)T;
    a.n_taps = N;
    a.n_groups = (int)((c_count + 127) / 128);
    a.n_tiles = (int)((T + 127) / 128);
    const int64_t total = (int64_t)a.n_groups * a.n_tiles;
    const int grid = (int)std::min<int64_t>(p->sm_count, total);
    st = raise_max_smem((const void*)zg_fir_tc_kernel, p->opts.device, zgk::kTcSmemBytes);
    if (st != ZG_OK) return st;
    a.split_mode = tune_env("ZG_TUNE_FIR_SPLIT");
    const bool prof = tune_env("ZG_TUNE_FIR_PROF") != 0;          // tuning only: where the roles of a CTA wait
    if (prof) ZG_CUDA(cudaMalloc(&a.prof, (size_t)grid * 8 * sizeof(long long)));
    void* args[] = {&a};
    ZG_CUDA(cudaLaunchKernel((const void*)zg_fir_tc_kernel, dim3(grid), dim3(zgk::kTcThreads), args, zgk::kTcSmemBytes, stream));
    if (prof) {
        std::vector<long long> h((size_t)grid * 8);
        ZG_CUDA(cudaMemcpy(h.data(), a.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.prof);
        double s[8] = {};
        for (int b = 0; b < grid; ++b)
            for (int k = 0; k < 8; ++k) s[k] += (double)h[(size_t)b * 8 + k] / grid;
        std::fprintf(stderr, "[zg_fir_tc] per CTA: %.0f blocks in %.0f cycles (%.0f / block); MMA thread waits: TMA %.0f, split %.0f, "
                             "accumulator drain %.0f; producer waits for a free stage %.0f; epilogue waits for an accumulator %.0f, drains %.0f\n",
                     s[4], s[0], s[0] / std::max(s[4], 1.0), s[1], s[2], s[3], s[5], s[6], s[7]);
    }
    // the delay line after this block (ping-pong: a later launch reads what this one writes)
    zgk::zg_fir_state_kernel<<<dim3((unsigned)((c_count + 31) / 32), (unsigned)((N - 1 + 31) / 32)), 256, 0, stream>>>(
        static_cast<const float*>(in[0]), ld_in, p->d_state + c_begin, p->d_state_alt + c_begin, p->ch_stride, (int)c_count,
        (int)T, N - 1);
    ZG_CUDA(cudaGetLastError());
    if (advance) {
        std::swap(p->d_state, p->d_state_alt);
        p->stream_pos += T;
    }
    p->launches += 2;
    p->last_smem = zgk::kTcSmemBytes;
    p->last_threads = zgk::kTcThreads;
    p->last_stages = zgk::kTcStages;
    p->last_boxes = (int)((total + grid - 1) / grid);
    p->last_grid = grid;
    return ZG_OK;
}

int launch_fir(zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t ld_in, int64_t ld_out,
               cudaStream_t stream, int64_t c_begin, int64_t c_count, bool advance) {
    if (p->fir_tc && T >= 256 && T <= (1ll << 30)) return launch_fir_tc(p, in, out, T, ld_in, ld_out, stream, c_begin, c_count, advance);
    const int N = (int)p->fir.taps.size();
    const int H = (N - 1 + 31) / 32;
    const int n_taps_pad = (N + 15) / 16 * 16;
    const int fixed = 1024 /*alignment slack*/ + n_taps_pad * 4 + 64 /*barriers*/ + H * zgk::kTileBytes;
    int W = (p->max_smem_optin - fixed) / (4 * zgk::kTileBytes);
    int want = 8;
    if (int w = tune_env("ZG_TUNE_WPC")) want = w;
    W = std::min({W, want, 16});
    if (W < 1) return fail(ZG_ERR_UNSUPPORTED, "FIR has too many taps for the shared-memory ring");
    const int64_t total_boxes = (T + zgk::kTileT - 1) / zgk::kTileT;
    W = (int)std::max<int64_t>(1, std::min<int64_t>(W, total_boxes));
    const int NR = H + 2 * W;
    const int64_t groups = (c_count + 31) / 32;
    int64_t n_segs = std::max<int64_t>(1, (2 * (int64_t)p->sm_count + groups - 1) / groups);
    if (int sg = tune_env("ZG_TUNE_SEGS")) n_segs = sg;
    const int64_t min_seg = ((std::max(H, 4 * W) + W - 1) / W) * W;          // >= H, whole steps, worth a prologue
    int64_t seg_boxes = (total_boxes + n_segs - 1) / n_segs;
    seg_boxes = std::max<int64_t>(min_seg, (seg_boxes + W - 1) / W * W);
    n_segs = (total_boxes + seg_boxes - 1) / seg_boxes;
    if (groups * n_segs > 0x7fffffffLL) return fail(ZG_ERR_ARG, "block too large for one launch");

    zgk::FirArgs a;
    std::memset(&a, 0, sizeof a);
    int st = encode_map(p, &a.in_map, in[0], c_count, T, ld_in, 32);
    if (st != ZG_OK) return st;
    st = encode_map(p, &a.out_map, out[0], c_count, T, ld_out, 32);
    if (st != ZG_OK) return st;
    a.state_in = p->d_state + c_begin;
    a.state_out = p->d_state_alt + c_begin;
    a.taps = p->d_taps;
    a.ch_stride = p->ch_stride;
    a.channels = (int)c_count;
    a.n_samples = (int)T;
    a.n_taps = N;
    a.hist_boxes = H;
    a.ring_boxes = NR;
    a.seg_boxes = (int)seg_boxes;
    a.n_segs = (int)n_segs;
    const int smem = fixed + (NR - H + 2 * W) * zgk::kTileBytes;
    const void* fn = fir_kernel_for(p->exact, p->interleaved);
    st = raise_max_smem(fn, p->opts.device, smem);
    if (st != ZG_OK) return st;
    void* args[] = {&a};
    ZG_CUDA(cudaLaunchKernel(fn, dim3((unsigned)(groups * n_segs)), dim3(W * 32), args, smem, stream));
    if (advance) {
        std::swap(p->d_state, p->d_state_alt);      // stream order: the next launch reads what this one wrote
        p->stream_pos += T;
    }
    p->launches += 1;
    p->last_smem = smem;
    p->last_threads = W * 32;
    p->last_stages = NR;
    p->last_boxes = (int)seg_boxes;
    p->last_grid = (int)(groups * n_segs);
    return ZG_OK;
}

// Long delay lines are rings in d_state (row t mod depth holds the value pushed at tick t); the ABI shows every
// line oldest value first, like the reference's std::array after rotate_push_back (flowz.hpp:130-148).  Slot j of
// a line of depth D holds the value pushed D - j ticks ago = ring row (stream_pos + j) mod D.
void rotate_rings(const zg_plan* p, float* host /* [n_state][C] */, bool ring_to_abi) {
    for (int l : p->ring.out_lines) {
        const IrLine& line = p->ir.lines[l];
        const int D = line.depth;
        const int64_t phase = ((p->stream_pos % D) + D) % D;
        std::vector<float> tmp((size_t)D * p->C);
        float* rows = host + (size_t)line.offset * p->C;
        for (int j = 0; j < D; ++j) {
            const int ring_row = (int)((phase + j) % D);
            const float* src = rows + (size_t)(ring_to_abi ? ring_row : j) * p->C;
            float* dst = tmp.data() + (size_t)(ring_to_abi ? j : ring_row) * p->C;
            std::memcpy(dst, src, p->C * sizeof(float));
        }
        std::memcpy(rows, tmp.data(), tmp.size() * sizeof(float));
    }
}

int check_io(const zg_plan* p, const void* const* in, void* const* out, int64_t T, int64_t ld_in, int64_t ld_out) {
    if (T < 0) return fail(ZG_ERR_ARG, "n_samples < 0");
    if (T > 0x7fffffffLL) return fail(ZG_ERR_ARG, "n_samples too large for one block");
    const int64_t min_ld = p->interleaved ? p->C : T;
    const int ldm = 16 / p->io;                  // rows must start 16-byte aligned (TMA global strides)
    if (p->n_buf_in > 0 && (ld_in < min_ld || ld_in % ldm)) return fail(ZG_ERR_ARG, "ld_in too small or not a multiple of 16 bytes");
    if (ld_out < min_ld || ld_out % ldm) return fail(ZG_ERR_ARG, "ld_out too small or not a multiple of 16 bytes");
    if (p->n_buf_in > 0 && !in) return fail(ZG_ERR_ARG, "in is NULL");
    if (!out) return fail(ZG_ERR_ARG, "out is NULL");
    for (int k = 0; k < p->ir.n_in; ++k) {
        if (p->synth_mask & (1u << k)) continue;
        if (!in[k] || ((uintptr_t)in[k] & 15)) return fail(ZG_ERR_ARG, "input buffer NULL or not 16-byte aligned");
    }
    for (int o = 0; o < p->ir.n_out; ++o)
        if (!out[o] || ((uintptr_t)out[o] & 15)) return fail(ZG_ERR_ARG, "output buffer NULL or not 16-byte aligned");
    // In place: an output may BE an input (same pointer, same pitch) -- every tile of every input is in shared memory
    // before any output of that tile is stored, and a tile is stored exactly where it was loaded.  Any other overlap
    // is refused, and so is aliasing for the FIR kernel (later time segments read their history from the input).
    const int64_t rows = p->interleaved ? T : p->C, cols = p->interleaved ? p->C : T;
    auto extent = [&](int64_t ld) { return rows > 0 ? ((rows - 1) * ld + cols) * (int64_t)p->io : 0; };
    auto overlap = [](const void* a, int64_t na, const void* b, int64_t nb) {
        const uintptr_t x = (uintptr_t)a, y = (uintptr_t)b;
        return x < y + (uintptr_t)nb && y < x + (uintptr_t)na;
    };
    for (int o = 0; o < p->ir.n_out; ++o) {
        for (int q = 0; q < o; ++q)
            if (overlap(out[o], extent(ld_out), out[q], extent(ld_out))) return fail(ZG_ERR_ARG, "output buffers overlap");
        for (int k = 0; k < p->ir.n_in; ++k) {
            if ((p->synth_mask & (1u << k)) || !overlap(out[o], extent(ld_out), in[k], extent(ld_in))) continue;
            if (out[o] != in[k] || ld_in != ld_out)
                return fail(ZG_ERR_ARG, "an output overlaps an input without being the same buffer (same pointer and pitch)");
            if (p->is_fir) return fail(ZG_ERR_UNSUPPORTED, "the FIR kernel does not run in place (its history is read from the input block)");
        }
    }
    return ZG_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------

extern "C" {

int zg_graph_kernel_compile(const zg_graph* g, const zg_plan_opts* opts, int uniform_params, int want_cubin,
                            char* buf, size_t capacity, size_t* size) {
    if (!g || !opts || !size) return fail(ZG_ERR_ARG, "NULL argument");
    const Ir& ir = g->ir_f32;
    if (!ir.all_f32()) return fail(ZG_ERR_UNSUPPORTED, "the device path evaluates fp32 graphs only");
    if (opts->io_dtype != ZG_F32 && opts->io_dtype != ZG_BF16) return fail(ZG_ERR_ARG, "io_dtype must be ZG_F32 or ZG_BF16");
    RingPlan ring;
    int ring_pf = 1;
    const Ir kir = split_for_kernel(ir, opts->layout == ZG_INTERLEAVED, opts->io_dtype == ZG_BF16 ? 2 : 4, ring, ring_pf);
    if (kir.n_state > zgk::kMaxState || (int)ring.taps.size() > zgk::kMaxRingIn || (int)ring.out_lines.size() > zgk::kMaxRingOut)
        return fail(ZG_ERR_UNSUPPORTED, "too much delay state for the generated kernel");
    unsigned synth = 0;
    for (int k = 0; k < ir.n_in; ++k)
        if (opts->input_kind[k] != ZG_IN_BUFFER) synth |= 1u << k;
    const bool exact = opts->mode == ZG_MODE_EXACT;
    std::string text = jit_source(kir, exact, opts->layout == ZG_INTERLEAVED, uniform_params != 0, synth,
                                  opts->io_dtype == ZG_BF16 ? 2 : 4, (int)ring.taps.size(), (int)ring.out_lines.size(), ring_pf);
    std::vector<char> cubin;
    const char* data = text.data();
    size_t n = text.size();
    if (want_cubin) {
        int st = jit_cubin(text, exact, cubin);
        if (st != ZG_OK) return st;
        data = cubin.data();
        n = cubin.size();
    }
    *size = n;
    if (buf) {
        if (capacity < n) return fail(ZG_ERR_ARG, "buffer too small");
        std::memcpy(buf, data, n);
    }
    return ZG_OK;
}

void zg_plan_opts_default(zg_plan_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->device = 0;
    o->channels = 1;
    o->mode = ZG_MODE_EXACT;
    o->layout = ZG_PLANAR;
    o->io_dtype = ZG_F32;
    o->lanes_per_channel = 0;
}

int zg_plan_create(const zg_graph* g, const zg_plan_opts* opts, zg_plan** out) {
    NvtxRange range("zg_plan_create");
    if (!g || !opts || !out) return fail(ZG_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (opts->channels < 1 || opts->channels > 0x7fffffe0LL) return fail(ZG_ERR_ARG, "channels out of range");
    if (opts->io_dtype != ZG_F32 && opts->io_dtype != ZG_BF16) return fail(ZG_ERR_ARG, "io_dtype must be ZG_F32 or ZG_BF16");
    const bool bf16 = opts->io_dtype == ZG_BF16;
    if (opts->mode != ZG_MODE_EXACT && opts->mode != ZG_MODE_FAST) return fail(ZG_ERR_ARG, "bad mode");
    if (opts->layout != ZG_PLANAR && opts->layout != ZG_INTERLEAVED) return fail(ZG_ERR_ARG, "bad layout");
    if (opts->time_parallel < ZG_TP_AUTO || opts->time_parallel > ZG_TP_TWO_PASS) return fail(ZG_ERR_ARG, "bad time_parallel");
    if (opts->fir_tensor_cores < 0 || opts->fir_tensor_cores > 1) return fail(ZG_ERR_ARG, "bad fir_tensor_cores");
    if (opts->section_warps < 0 || opts->section_warps > 2) return fail(ZG_ERR_ARG, "bad section_warps");
    const Ir& ir = g->ir_f32;
    if (!ir.all_f32())
        return fail(ZG_ERR_UNSUPPORTED,
                    "the device path evaluates fp32 graphs only; this graph has int or double terminals "
                    "(tick it on the host with zg_voice_tick)");
    if (ir.n_out < 1) return fail(ZG_ERR_UNSUPPORTED, "graph has no outputs");
    FirMatch fir;
    bool any_synth = false;
    for (int k = 0; k < ir.n_in; ++k) any_synth = any_synth || opts->input_kind[k] != ZG_IN_BUFFER;
    const bool is_fir = !opts->force_jit && !any_synth && !bf16 && match_fir(ir, fir);
    // generated kernel: delay lines deeper than kRegLineDepth live in HBM as rings (zg_ir.hpp); what must fit in
    // registers is the rest -- short lines and the near windows of the long ones
    RingPlan ring;
    int ring_pf = 1;
    const Ir kir = split_for_kernel(ir, opts->layout == ZG_INTERLEAVED, bf16 ? 2 : 4, ring, ring_pf);
    if (!is_fir && (kir.n_state > zgk::kMaxState || (int)ring.taps.size() > zgk::kMaxRingIn ||
                    (int)ring.out_lines.size() > zgk::kMaxRingOut || ir.n_in + ring.taps.size() > 32))
        return fail(ZG_ERR_UNSUPPORTED,
                    "graph keeps " + std::to_string(kir.n_state) + " floats of register-resident delay state per channel (limit " +
                        std::to_string(zgk::kMaxState) + "), " + std::to_string(ring.out_lines.size()) + " long delay lines (limit " +
                        std::to_string(zgk::kMaxRingOut) + ") with " + std::to_string(ring.taps.size()) + " far reads (limit " +
                        std::to_string(zgk::kMaxRingIn) +
                        "); a dense FIR c0*_1 + c1*_1[_1] + ... on the input (planar fp32) has its own kernel");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(ZG_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                     " (zignal-b200 has no CPU fallback for block evaluation)");
    }
    if (opts->device < 0 || opts->device >= ndev) return fail(ZG_ERR_ARG, "bad device ordinal");
    ZG_ON_DEVICE(opts->device);
    ZG_CUDA(cudaFree(nullptr));
    cudaDeviceProp prop;
    ZG_CUDA(cudaGetDeviceProperties(&prop, opts->device));
    if (prop.major != 10)
        return fail(ZG_ERR_CUDA, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) +
                                     std::to_string(prop.minor) + "); this library contains sm_100a code only");
    if (!driver().ok) return fail(ZG_ERR_CUDA, driver().why);

    auto p = std::make_unique<zg_plan>();
    p->ir = ir;
    p->opts = *opts;
    p->C = opts->channels;
    p->ch_stride = (p->C + 31) / 32 * 32;
    p->exact = opts->mode == ZG_MODE_EXACT;
    p->interleaved = opts->layout == ZG_INTERLEAVED;
    p->io = bf16 ? 2 : 4;
    p->sm_count = prop.multiProcessorCount;
    p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (int k = 0; k < ir.n_in; ++k) {
        int kind = opts->input_kind[k];
        if (kind == ZG_IN_DIRAC) { p->synth_mask |= 1u << k; p->dirac_mask |= 1u << k; }
        else if (kind == ZG_IN_ZERO) p->synth_mask |= 1u << k;
        else if (kind != ZG_IN_BUFFER) return fail(ZG_ERR_ARG, "bad input_kind");
        else p->n_buf_in += 1;
    }

    p->is_fir = is_fir;
    if (is_fir) p->fir = fir;
    // bf16 storage runs the generated kernel (the prebuilt ones are instantiated for fp32 storage only;
    // the generated tick of a biquad cascade is the same arithmetic, tests/test_gpu_parity.py)
    p->is_biquad = !is_fir && !bf16 && match_df1_cascade(ir, p->bq) && p->synth_mask == 0;
    const int want_lanes = tune_env("ZG_TUNE_LANES") ? tune_env("ZG_TUNE_LANES") : opts->lanes_per_channel;
    if (want_lanes < 0 || (want_lanes > 1 && !(p->is_biquad && !opts->force_jit && !p->interleaved &&
                                               want_lanes == p->bq.sections && (want_lanes == 2 || want_lanes == 4))))
        return fail(ZG_ERR_UNSUPPORTED,
                    "lanes_per_channel > 1 needs a planar biquad cascade of exactly that many (2 or 4) sections");
    if (p->is_biquad && !opts->force_jit && !p->interleaved && (p->bq.sections == 2 || p->bq.sections == 4)) {
        // one lane per channel leaves SM schedulers idle below ~2 warps per scheduler: spread the sections
        // measured crossover (4 sections, T >= 16k; profiles/r01_sweep_ns_lanes.jsonl): 8192 channels K1b 0.60 ms
        // vs K1 1.18 ms, 16 384 channels K1b 0.79 vs K1 0.58 -> K1b below ~2.5 warps of channels per SM
        const bool too_few = (p->C + 31) / 32 < (int64_t)p->sm_count * 5 / 2;
        if (want_lanes > 1 || (want_lanes == 0 && too_few)) p->lanes = p->bq.sections;
    }
    p->lanes_now = p->lanes;
    // time segments: FAST mode (they re-associate the arithmetic in time), linear or affine tick, all delay lines in
    // registers, one lane per channel not ruled out by an explicit lanes_per_channel
    p->linearity = ir_linearity(ir);
    p->scan_ok = !is_fir && !p->exact && p->linearity != ZG_NONLINEAR && !ring.any() && want_lanes <= 1 && ir.n_state <= zgk::kMaxState;
    if (opts->time_parallel >= ZG_TP_WARMUP && !p->scan_ok)
        return fail(ZG_ERR_UNSUPPORTED,
                    "time_parallel needs ZG_MODE_FAST, a linear or affine tick (zg_graph_linearity), delay lines of at most " +
                        std::to_string(kRegLineDepth) + " samples, lanes_per_channel <= 1 and a graph that is not a dense FIR");
    if (p->is_fir) {
        p->kernel_n_state = ir.n_state;
        p->kernel_n_param = 0;
        // FAST mode, planar, 2..256 taps: the Toeplitz contraction on tensor cores (3xTF32); EXACT keeps the CUDA-core
        // kernel, whose left-to-right sum is bit-identical to the reference (ZG_TUNE_FIR_TC=1 / 2 forces off / on)
        p->fir_tc = !p->exact && !p->interleaved && fir.taps.size() >= 2 && fir.taps.size() <= 256 && opts->fir_tensor_cores != 1 &&
                    tune_env("ZG_TUNE_FIR_TC") != 1;
        p->kernel_name = p->fir_tc ? "zg_fir_tc<" + std::to_string(fir.taps.size()) + " taps,3xtf32,planar>"
                                   : "zg_fir<" + std::to_string(fir.taps.size()) + (p->exact ? " taps,exact," : " taps,fma,") +
                                         (p->interleaved ? "interleaved>" : "planar>");
    } else if (p->is_biquad && !opts->force_jit) {
        const int S = p->bq.sections;
        p->kernel_n_state = 2 * (S + 1);
        p->kernel_n_param = 5 * S;
        p->state_row.resize(p->kernel_n_state);
        for (int k = 0; k <= S; ++k) {
            const IrLine& l = ir.lines[p->bq.signal_line[k]];
            p->state_row[2 * k] = l.offset;          // two ticks ago
            p->state_row[2 * k + 1] = l.offset + 1;  // one tick ago
        }
        char nm[96];
        if (p->lanes > 1)
            std::snprintf(nm, sizeof nm, "zg_biquad_df1_lanes<%d,%s,planar>", S, p->exact ? "exact" : "fma");
        else
            std::snprintf(nm, sizeof nm, "zg_biquad_df1<%d,%s,%s>", S, p->exact ? "exact" : "fma",
                          p->interleaved ? "interleaved" : "planar");
        p->kernel_name = nm;
    } else {
        p->kir = kir;
        p->ring = ring;
        p->ring_pf = ring_pf;
        p->kernel_n_state = kir.n_state;
        p->kernel_n_param = ir.n_params;
        p->state_row.assign(kir.n_state, 0);       // filled per launch (window slots move with the stream position)
        p->kernel_name = std::string("zg_graph_kernel<jit,") + (p->exact ? "exact," : "fma,") +
                         (p->interleaved ? "interleaved" : "planar") + (bf16 ? ",bf16>" : ">");
    }

    {
        int arith = 0, mul = 0;
        for (const IrNode& n : ir.nodes) {
            if (n.op == IrOp::Add || n.op == IrOp::Sub || n.op == IrOp::Mul || n.op == IrOp::Div || n.op == IrOp::Neg) ++arith;
            if (n.op == IrOp::Mul) ++mul;
        }
        // FAST: a product feeding a sum contracts into one FMA (at most one product per sum)
        p->tick_ops = p->exact ? arith : arith - std::min(mul, arith - mul);
        const int bytes = p->io * std::max(1, p->n_buf_in + ir.n_out);
        p->light_tick = 2 * p->tick_ops < 7 * bytes;
    }
    const size_t state_floats = (size_t)std::max(ir.n_state, 1) * p->ch_stride;
    ZG_CUDA(cudaMalloc(&p->d_state, state_floats * sizeof(float)));
    ZG_CUDA(cudaMemset(p->d_state, 0, state_floats * sizeof(float)));
    if (p->is_fir) {
        ZG_CUDA(cudaMalloc(&p->d_state_alt, state_floats * sizeof(float)));
        ZG_CUDA(cudaMemset(p->d_state_alt, 0, state_floats * sizeof(float)));
    }
    p->h_params.assign(ir.n_params, std::vector<float>(1, 0.f));
    p->param_on_device.assign(ir.n_params, 0);
    p->params_dirty = true;

    // compile / pick the kernel now so that plan creation is where build errors surface
    int st = sync_params(p.get());
    if (st != ZG_OK) return st;
    Variant* v = nullptr;
    st = get_variant(p.get(), p->uniform_now, v);
    if (st != ZG_OK) return st;
    if (p->scan_ok && opts->time_parallel >= ZG_TP_WARMUP) {
        // a plan that asks for time segments gets the kernel built with them now, not at its first launch
        p->seg_now = true;
        p->lanes_now = 1;
        st = get_variant(p.get(), p->uniform_now, v);
        p->seg_now = false;
        p->lanes_now = p->lanes;
        if (st != ZG_OK) return st;
    }
    *out = p.release();
    return ZG_OK;
}

void zg_plan_destroy(zg_plan* p) { delete p; }

int zg_split_plan_query(int sections, int exact, int64_t channels, int64_t samples, int sm_count, int max_smem,
                        int segments, int warmup_samples, zg_split_plan* out) {
    if (!out || sections < 1 || sections > 8 || channels < 1 || samples < 1 || sm_count < 1 || max_smem < 1) return 0;
    // the planner reads a plan: one with just the fields it looks at (no device is touched)
    std::unique_ptr<zg_plan> p(new zg_plan());
    p->opts.device = -1;
    p->is_biquad = true;
    p->exact = exact != 0;
    p->bq.sections = sections;
    p->C = channels;
    p->ch_stride = (channels + 31) / 32 * 32;
    p->sm_count = sm_count;
    p->max_smem_optin = max_smem;
    p->uniform_now = true;
    p->kernel_n_state = p->ir.n_state = 2 * (sections + 1);
    // the lanes the auto rule of zg_plan_create gives a cascade of 2 or 4 sections on few channels
    p->lanes = (sections == 2 || sections == 4) && (channels + 31) / 32 < (int64_t)sm_count * 5 / 2 ? sections : 1;
    p->lanes_now = 1;
    Segments sg;
    if (segments > 1) {
        sg.mode = 1;
        sg.n = segments;
        sg.warm = warmup_samples;
    }
    SplitGeometry g{};
    if (!choose_split(p.get(), samples, channels, g, segments > 1 ? &sg : nullptr)) return 0;
    out->groups_per_cta = g.groups;
    out->warps_per_group = g.wpg;
    out->sections_per_warp = g.spw;
    out->grid = g.grid;
    out->threads_per_cta = g.groups * g.wpg * 32;
    out->stages = g.stages;
    out->boxes_per_tile = g.boxes;
    out->boxes_per_handover = g.hand_boxes;
    out->smem_bytes = g.smem;
    out->segments = g.n_segs;
    out->segment_boxes = g.seg_boxes;
    out->warmup_boxes = g.warm_boxes;
    return 1;
}

int zg_plan_get_info(const zg_plan* p, zg_plan_info* info) {
    if (!p || !info) return fail(ZG_ERR_ARG, "NULL argument");
    std::memset(info, 0, sizeof *info);
    std::string name = p->kernel_name;
    if (p->lanes > 1 && p->lanes_now == 1)                      // a K1b plan whose last launch was cut in time instead
        name = "zg_biquad_df1<" + std::to_string(p->bq.sections) + (p->exact ? ",exact,planar>" : ",fma,planar>");
    if (p->split_now)
        name = "zg_biquad_df1_split<" + std::to_string(p->bq.sections) + (p->exact ? ",exact,planar," : ",fma,planar,") +
               std::to_string(p->bq.sections / std::max(p->split_spw, 1)) + " warps per group" +
               (p->split_hand_boxes > 1 ? "," + std::to_string(p->split_hand_boxes) + " boxes per hand-over>" : ">");
    std::snprintf(info->kernel, sizeof info->kernel, "%s%s%s", name.c_str(),
                  variant_is_sym(p) && !(p->split_now && p->split_hand_boxes > 1) ? "+b0=b2" : "",           // the product-reusing tick (kernels/zg_biquad.cuh)
                  p->last_seg_mode == 1 ? "+segments:warm-up" : p->last_seg_mode == 2 ? "+segments:two-pass" : "");
    const Variant& v = p->variant[variant_index(p)];
    info->jit = (v.prebuilt || p->is_fir) ? 0 : 1;
    info->lanes_per_channel = p->lanes_now;
    info->time_segments = p->last_segs;
    info->segment_samples = p->last_seg_len;
    info->warmup_samples = p->last_seg_warm;
    info->linearity = p->linearity;
    info->host_chunks = p->last_host_chunks;
    info->regs_per_thread = p->split_now ? p->split_regs : v.regs;
    info->smem_bytes = p->last_smem;
    info->launches = p->launches;
    info->threads_per_cta = p->last_threads;
    info->stages = p->last_stages;
    info->boxes = p->last_boxes;
    info->uniform_params = p->uniform_now ? 1 : 0;
    return ZG_OK;
}

int zg_process(zg_plan* p, const void* const* in, void* const* out, int64_t n_samples, int64_t ld_in,
               int64_t ld_out, void* stream) {
    if (!p) return fail(ZG_ERR_ARG, "plan is NULL");
    NvtxRange range("zg_process");
    int st = check_io(p, in, out, n_samples, ld_in, ld_out);
    if (st != ZG_OK) return st;
    if (n_samples == 0) return ZG_OK;
    ZG_ON_DEVICE(p->opts.device);
    return launch(p, in, out, n_samples, ld_in, ld_out, (cudaStream_t)stream, 0, p->C, true);
}

int zg_process_host(zg_plan* p, const void* const* in, void* const* out, int64_t n_samples, int64_t ld_in,
                    int64_t ld_out) {
    if (!p) return fail(ZG_ERR_ARG, "plan is NULL");
    if (n_samples < 0) return fail(ZG_ERR_ARG, "n_samples < 0");
    if (n_samples == 0) return ZG_OK;
    NvtxRange range("zg_process_host");
    ZG_ON_DEVICE(p->opts.device);
    if (!p->own_stream) ZG_CUDA(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    if (!p->h2d_stream) ZG_CUDA(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
    if (!p->d2h_stream) ZG_CUDA(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
    // device staging: same shape as the host buffers, rows padded to a multiple of 4 floats
    const int64_t rows = p->interleaved ? n_samples : p->C;
    const int64_t cols = p->interleaved ? p->C : n_samples;
    const int es = p->io;                                   // bytes per sample
    const int ldm = 16 / es;
    const int64_t ld = (cols + ldm - 1) / ldm * ldm;
    if ((p->n_buf_in > 0 && ld_in < cols) || ld_out < cols) return fail(ZG_ERR_ARG, "leading dimension too small");
    const size_t per_buf = (size_t)rows * ld * es;          // bytes
    const size_t need = per_buf * (p->n_buf_in + p->ir.n_out);
    if (need > p->d_stage_bytes) {
        if (p->d_stage) cudaFree(p->d_stage);
        p->d_stage = nullptr;
        p->d_stage_bytes = 0;
        ZG_CUDA(cudaMalloc(&p->d_stage, need));
        p->d_stage_bytes = need;
    }
    unsigned char* d_in[ZG_MAX_WIRES] = {};
    unsigned char* d_out[ZG_MAX_WIRES] = {};
    size_t slot = 0;
    for (int k = 0; k < p->ir.n_in; ++k) {
        if (p->synth_mask & (1u << k)) continue;
        if (!in || !in[k]) return fail(ZG_ERR_ARG, "input buffer is NULL");
        d_in[k] = p->d_stage + slot++ * per_buf;
    }
    for (int o = 0; o < p->ir.n_out; ++o) {
        if (!out || !out[o]) return fail(ZG_ERR_ARG, "output buffer is NULL");
        d_out[o] = p->d_stage + slot++ * per_buf;
    }
    {
        const void* ci[ZG_MAX_WIRES];
        void* co[ZG_MAX_WIRES];
        for (int k = 0; k < ZG_MAX_WIRES; ++k) { ci[k] = d_in[k]; co[k] = d_out[k]; }
        int st = check_io(p, ci, co, n_samples, ld, ld);
        if (st != ZG_OK) return st;
    }

    // The block is cut into row chunks (planar: channel ranges, independent; interleaved: time ranges,
    // launched in order, state carried in HBM) and streamed: H2D of chunk k+1, the kernel of chunk k
    // and D2H of chunk k-1 overlap on three streams, so a PCIe-bound call costs one direction, not two.
    const int64_t row_bytes = cols * es * std::max(1, std::max(p->n_buf_in, p->ir.n_out));
    // (the first H2D and the last D2H overlap nothing: 1/n_chunks of each direction is exposed, so many
    // chunks -- but each still tens of MB, far above the per-copy and per-launch overheads)
    int64_t n_chunks = std::min<int64_t>(32, (rows * row_bytes) / (32ll << 20));
    if (int c = tune_env("ZG_TUNE_HOST_CHUNKS")) n_chunks = c;
    n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, rows / 32));
    int64_t chunk_rows = (rows + n_chunks - 1) / n_chunks;
    chunk_rows = (chunk_rows + 31) / 32 * 32;
    n_chunks = (rows + chunk_rows - 1) / chunk_rows;
    while ((int64_t)p->events.size() < 2 * n_chunks) {
        cudaEvent_t e;
        ZG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->events.push_back(e);
    }
    for (int64_t c = 0; c < n_chunks; ++c) {
        NvtxRange chunk_range("zg_process_host: chunk (H2D, kernel, D2H enqueued)");
        const int64_t r0 = c * chunk_rows, nr = std::min(chunk_rows, rows - r0);
        for (int k = 0; k < p->ir.n_in; ++k) {
            if (!d_in[k]) continue;
            ZG_CUDA(cudaMemcpy2DAsync(d_in[k] + r0 * ld * es, ld * es, (const unsigned char*)in[k] + r0 * ld_in * es,
                                      ld_in * es, cols * es, nr, cudaMemcpyHostToDevice, p->h2d_stream));
        }
        ZG_CUDA(cudaEventRecord(p->events[2 * c], p->h2d_stream));
        ZG_CUDA(cudaStreamWaitEvent(p->own_stream, p->events[2 * c], 0));
        const void* ci[ZG_MAX_WIRES] = {};
        void* co[ZG_MAX_WIRES] = {};
        for (int k = 0; k < ZG_MAX_WIRES; ++k) {
            ci[k] = d_in[k] ? d_in[k] + r0 * ld * es : nullptr;
            co[k] = d_out[k] ? d_out[k] + r0 * ld * es : nullptr;
        }
        int st = p->interleaved ? launch(p, ci, co, nr, ld, ld, p->own_stream, 0, p->C, true)
                                : launch(p, ci, co, n_samples, ld, ld, p->own_stream, r0, nr, c + 1 == n_chunks);
        if (st != ZG_OK) {
            cudaDeviceSynchronize();
            return st;
        }
        ZG_CUDA(cudaEventRecord(p->events[2 * c + 1], p->own_stream));
        ZG_CUDA(cudaStreamWaitEvent(p->d2h_stream, p->events[2 * c + 1], 0));
        for (int o = 0; o < p->ir.n_out; ++o)
            ZG_CUDA(cudaMemcpy2DAsync((unsigned char*)out[o] + r0 * ld_out * es, ld_out * es, d_out[o] + r0 * ld * es, ld * es,
                                      cols * es, nr, cudaMemcpyDeviceToHost, p->d2h_stream));
    }
    p->last_host_chunks = (int)n_chunks;
    ZG_CUDA(cudaStreamSynchronize(p->d2h_stream));
    ZG_CUDA(cudaStreamSynchronize(p->own_stream));
    return ZG_OK;
}

int zg_state_reset(zg_plan* p) {
    if (!p) return fail(ZG_ERR_ARG, "plan is NULL");
    ZG_ON_DEVICE(p->opts.device);
    ZG_CUDA(cudaDeviceSynchronize());
    ZG_CUDA(cudaMemset(p->d_state, 0, (size_t)std::max(p->ir.n_state, 1) * p->ch_stride * sizeof(float)));
    p->stream_pos = 0;
    return ZG_OK;
}

int zg_state_get(zg_plan* p, float* host, size_t n_floats) {
    if (!p || !host) return fail(ZG_ERR_ARG, "NULL argument");
    if (n_floats != (size_t)p->ir.n_state * p->C) return fail(ZG_ERR_ARG, "expected n_state * channels floats");
    if (n_floats == 0) return ZG_OK;
    ZG_ON_DEVICE(p->opts.device);
    ZG_CUDA(cudaDeviceSynchronize());
    ZG_CUDA(cudaMemcpy2D(host, p->C * 4, p->d_state, p->ch_stride * 4, p->C * 4, p->ir.n_state, cudaMemcpyDeviceToHost));
    rotate_rings(p, host, true);
    return ZG_OK;
}

int zg_state_set(zg_plan* p, const float* host, size_t n_floats) {
    if (!p || !host) return fail(ZG_ERR_ARG, "NULL argument");
    if (n_floats != (size_t)p->ir.n_state * p->C) return fail(ZG_ERR_ARG, "expected n_state * channels floats");
    if (n_floats == 0) return ZG_OK;
    ZG_ON_DEVICE(p->opts.device);
    ZG_CUDA(cudaDeviceSynchronize());
    if (p->ring.any()) {
        std::vector<float> tmp(host, host + n_floats);
        rotate_rings(p, tmp.data(), false);
        ZG_CUDA(cudaMemcpy2D(p->d_state, p->ch_stride * 4, tmp.data(), p->C * 4, p->C * 4, p->ir.n_state, cudaMemcpyHostToDevice));
        return ZG_OK;
    }
    ZG_CUDA(cudaMemcpy2D(p->d_state, p->ch_stride * 4, host, p->C * 4, p->C * 4, p->ir.n_state, cudaMemcpyHostToDevice));
    return ZG_OK;
}

int zg_param_set(zg_plan* p, int index, const float* host_values, int64_t n) {
    if (!p || !host_values) return fail(ZG_ERR_ARG, "NULL argument");
    if (index < 0 || index >= (int)p->h_params.size()) return fail(ZG_ERR_ARG, "bad parameter index");
    if (n != 1 && n != p->C) return fail(ZG_ERR_ARG, "parameter needs 1 value or one per channel");
    ZG_ON_DEVICE(p->opts.device);
    ZG_CUDA(cudaDeviceSynchronize());      // a launch in flight may still be reading d_params
    p->h_params[index].assign(host_values, host_values + n);
    p->param_on_device[index] = 0;
    p->params_dirty = true;
    return ZG_OK;
}

int zg_param_set_device(zg_plan* p, int index, const float* device_values, int64_t n) {
    if (!p || !device_values) return fail(ZG_ERR_ARG, "NULL argument");
    if (index < 0 || index >= (int)p->h_params.size()) return fail(ZG_ERR_ARG, "bad parameter index");
    if (n != p->C) return fail(ZG_ERR_ARG, "a parameter set from device memory needs one value per channel");
    ZG_ON_DEVICE(p->opts.device);
    ZG_CUDA(cudaDeviceSynchronize());      // the producer of device_values and any launch still reading d_params
    if (!p->d_user_params) {
        const size_t bytes = p->h_params.size() * (size_t)p->ch_stride * sizeof(float);
        ZG_CUDA(cudaMalloc(&p->d_user_params, bytes));
        ZG_CUDA(cudaMemset(p->d_user_params, 0, bytes));
    }
    ZG_CUDA(cudaMemcpy(p->d_user_params + (size_t)index * p->ch_stride, device_values, p->C * sizeof(float),
                       cudaMemcpyDeviceToDevice));
    p->param_on_device[index] = 1;
    p->params_dirty = true;
    return ZG_OK;
}

}  // extern "C"
