/* zignal-b200 :: C ABI of the B200-native flowz evaluator.
 *
 * The reference (andre-bergner/zignal, /root/reference) is a header-only C++14 EDSL; it has no
 * FFI layer.  Its whole public surface for the evaluated path is
 *
 *     auto f = flowz::compile(expr);          flowz/flowz.hpp:1233-1249
 *     std::tuple<...> y = f(x1, ..., xN);     flowz/flowz.hpp:1225-1229   (one sample, one voice)
 *
 * This header is the boundary a binding for that path talks to.  The C++ shim
 * include/flowz/flowz.hpp (same spellings as the reference) and the Python test harness
 * (zignal_b200/__init__.py, ctypes) are the two clients.  Every entry point names the reference
 * code it stands in for.  All functions are extern "C", take plain pointers/sizes, never throw,
 * and return 0 (ZG_OK) or a negative zg_status; the message is in zg_last_error() (thread local).
 *
 * Expressions travel as text in flowz syntax with C++ operator precedence:
 *     _k              input wire k (1-based)                    flowz.hpp:75-82, 1252-1257
 *     _k[_n]  _k[-n]  wire k delayed by n samples               flowz.hpp:84-85 / prototypes
 *     _k<-n>  e[_n]   planned spellings: same; (e |= _1[_n])    TODO.md:8-9, 51-52
 *     a |= b, a >> b  series                                    flowz.hpp:92 / wires_mono_only.cpp:37
 *     a | b           parallel (splits the inputs)              flowz.hpp:91
 *     (a , b)         fan-out ("channel")                       flowz.hpp:90
 *     ~a              feedback                                  flowz.hpp:93
 *     + - * / unary - leaf arithmetic (C++ built-in semantics)  flowz.hpp:769-772
 *     0.5f 0.5 2      float / double / int literal terminals    flowz.hpp:68-72
 *     cplx{1,0}       std::complex<float> terminal (cplxd: double); type analysis only   test/tests.cpp:188,205
 *     $k              run-time parameter k (std::ref terminal)  flowz/README.md:42-63
 * Expression trees may be up to 1536 levels high (a 512-tap FIR written as one sum is 513); deeper text is refused
 * (ZG_ERR_PARSE / ZG_ERR_GRAPH) instead of exhausting the stack of the recursive analyses.
 */
#ifndef ZIGNAL_B200_H
#define ZIGNAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum zg_status {
    ZG_OK = 0,
    ZG_ERR_PARSE = -1,        /* malformed expression text */
    ZG_ERR_GRAPH = -2,        /* expression is not a valid flowz graph (a compile error in the reference) */
    ZG_ERR_ARG = -3,          /* bad argument */
    ZG_ERR_UNSUPPORTED = -4,  /* valid graph, but not supported by the requested path */
    ZG_ERR_CUDA = -5,         /* CUDA / NVRTC failure, or no usable device */
    ZG_ERR_INTERNAL = -6
} zg_status;

typedef enum zg_dtype {
    ZG_I32 = 0, ZG_F32 = 1, ZG_F64 = 2,
    ZG_BF16 = 3,            /* sample storage only */
    ZG_C64 = 4, ZG_C128 = 5,/* std::complex<float> / <double>: type analysis only (zg_expr_result_types) */
    ZG_TYPE_OPEN = -1       /* zg_expr_result_types: the reference's `absorber`, a type the feedback cycle leaves open */
} zg_dtype;

typedef struct zg_graph zg_graph; /* immutable, shareable between threads        */
typedef struct zg_voice zg_voice; /* one voice on the host: state_ of stateful_lambda */
typedef struct zg_plan zg_plan;   /* C channels of one graph on one GPU          */

const char* zg_last_error(void);
const char* zg_version(void);

/* ---- static analysis on a bare expression (no front panel) -----------------------------------
 * input_arity / output_arity            flowz.hpp:162-246
 * max_input_delays / min_input_delays   flowz.hpp:443-506  (which: 0 = max, 1 = min; -1 = unused)
 * make_canonical                        flowz.hpp:794-805  (text of the rewritten expression)   */
int zg_expr_arity(const char* expr, int* n_in, int* n_out);
int zg_expr_delays(const char* expr, int which, int* delays, int capacity, int* count);
int zg_expr_canonical(const char* expr, char* buf, size_t capacity);
/* ResultType                            flowz.hpp:515-644  (pinned by test/tests.cpp:182-232)
 * C++ type (zg_dtype) of every output wire when input wire k has type in_dtypes[k-1].  Fed-back wires start as
 * `absorber` and take the type of whatever they are combined with (:523-548); a wire no operand ever types is
 * reported as ZG_TYPE_OPEN (the reference's leftover absorber, :575-578).  *is_tuple = 0 where the reference
 * yields a bare scalar (terminal, delayed placeholder, plain arithmetic) and 1 where it yields a std::tuple.
 * cplx{re,im} / cplxd{re,im} spell std::complex<float> / <double> terminals; such graphs can be analysed,
 * not compiled (the reference's compile() keeps float state only, :1245).                        */
int zg_expr_result_types(const char* expr, const int* in_dtypes, int n_in, int* out_dtypes, int capacity,
                         int* count, int* is_tuple);

/* ---- compile()  flowz.hpp:1233-1249 -----------------------------------------------------------
 * front panel (:273-277) + canonical form (:794-935) + state layout (:685-725) + lowering to the
 * flat tick program that both the host tick and the CUDA kernels execute.
 * Graphs the reference compiles are walked exactly as it walks them.  Graphs it cannot compile because of a
 * feedback (nested loops, parallel combiners inside a loop: TODO.md:11-29) compile here as long as every loop
 * contains a delay: `~x` ties the first inputs of x to x's own outputs.  A loop without a delay, and a delayed
 * read the reference's state sizing does not cover (undefined behaviour there), are ZG_ERR_GRAPH.          */
int zg_graph_compile(const char* expr, zg_graph** out);
void zg_graph_destroy(zg_graph* g);

typedef struct zg_graph_info {
    int n_in;      /* input_arity of the user expression                                       */
    int n_out;     /* values one tick returns: output_arity, except where the reference's sequence passes surplus
                      inputs through uncounted (flowz.hpp:996-999 vs :238-247; `_1 |= (_1[_3] | _2[_1])`: 3, not 2) */
    int n_params;  /* number of $k parameters                                                  */
    int n_state;   /* floats of delay-line state per channel after line sharing                */
    int n_lines;   /* delay lines                                                              */
    int n_nodes;   /* SSA nodes of one tick                                                    */
    int all_f32;   /* 1 if the tick is pure fp32 when fed fp32 inputs (device path requirement) */
} zg_graph_info;
int zg_graph_get_info(const zg_graph* g, zg_graph_info* info);
/* canonical expression / SSA dump; valid until the graph is destroyed (print_state stand-in) */
const char* zg_graph_canonical(const zg_graph* g);
const char* zg_graph_dump(const zg_graph* g);
/* Which device kernel family the tick program is recognised as (planar layout, buffer inputs):
 *   "biquad_df1:<S>"  S direct-form-1 sections in series, test/benchmark.cpp:25-33  -> prebuilt K1 / K1b
 *   "fir:<N>"         c0*_1 + c1*_1[_1] + ... + c(N-1)*_1[_(N-1)], summed left to right -> prebuilt K3
 *   "generated"       any other fp32 graph: tick body generated from the SSA, NVRTC         -> K2
 *   "host-only"       int / double terminals: zg_voice_tick only
 * Pure host analysis (no device).                                                              */
int zg_graph_kernel_class(const zg_graph* g, char* buf, size_t capacity);
/* Is one tick a linear map of (inputs, delay-line state) -> (outputs, new state)?  Literals and $k parameters count as
 * coefficients.  ZG_LINEAR: superposition holds (every biquad / FIR / comb / oscillator graph of the reference's
 * benchmarks); ZG_AFFINE: linear plus constant terms (`_1 + 1`); ZG_NONLINEAR: a product or quotient of two signals.
 * Pure host analysis on the tick program -- what a time-parallel (scan) evaluation of few long channels would have
 * to ask first (the state-space idea is sketched in experimental_steps/tuprix.cpp:239-254). */
typedef enum zg_linearity { ZG_NONLINEAR = 0, ZG_AFFINE = 1, ZG_LINEAR = 2 } zg_linearity;
int zg_graph_linearity(const zg_graph* g, int* kind);
/* The state matrix A of a LINEAR / AFFINE tick, state' = A state + B x + c, for one set of $k values: row-major
 * [n_state][n_state] doubles in zg_state_get order, read off the tick program by unit-vector probes (every entry is
 * one coefficient path of the tick).  n_params must be the graph's parameter count.  Host analysis.              */
int zg_graph_state_matrix(const zg_graph* g, const float* params, int n_params, double* A, size_t capacity);
/* How fast the tick forgets its state: *K = the smallest multiple of `step` (<= k_max) with |A^K|_inf <= tol, 0 if
 * there is none (poles on or outside the unit circle).  ZG_TP_WARMUP below uses step = 4 boxes, k_max = 8192,
 * tol = 2^-30 and the worst channel.                                                                              */
int zg_graph_settling_time(const zg_graph* g, const float* params, int n_params, int step, int k_max, double tol, int* K);

/* ---- host voice: stateful_lambda (flowz.hpp:1181-1230) ----------------------------------------
 * zg_voice_tick is operator()(args...) for exactly n_in arguments; in_dtypes[i] says what C++
 * type argument i had (int stays int inside the tick, as with the reference's templates).
 * State starts at zero (:1191) and persists across ticks; zg_voice_clone copies it (:1207).    */
int zg_voice_create(const zg_graph* g, zg_voice** out);
int zg_voice_clone(const zg_voice* v, zg_voice** out);
void zg_voice_destroy(zg_voice* v);
int zg_voice_tick(zg_voice* v, const double* in, const int* in_dtypes, double* out, int* out_dtypes);
int zg_voice_set_param(zg_voice* v, int index, float value);
int zg_voice_state(zg_voice* v, float** state, int* n_state);

/* ---- device plan: the block evaluator (new; replaces the caller's per-sample for-loop,
 *      test/benchmark.cpp:137-147, for `channels` independent voices at once) ------------------ */
typedef enum zg_mode {
    ZG_MODE_EXACT = 0, /* separately rounded mul/add, lane per channel: bit-identical to the x86
                          non-FMA build of the reference (CMakeLists.txt:17-19)                 */
    ZG_MODE_FAST = 1   /* FMA contraction: rounds differently from the reference   */
} zg_mode;

typedef enum zg_layout {
    ZG_PLANAR = 0,      /* buffer[c * ld + t] : one contiguous block per channel */
    ZG_INTERLEAVED = 1  /* buffer[t * ld + c] : one frame per sample             */
} zg_layout;

typedef enum zg_input_kind {
    ZG_IN_BUFFER = 0, /* samples come from the caller's buffer                                  */
    ZG_IN_DIRAC = 1,  /* 1.0 at stream position 0, else 0; synthesised in the kernel (in[i] may be NULL) */
    ZG_IN_ZERO = 2    /* all zeros; synthesised in the kernel                                   */
} zg_input_kind;

/* Few, long channels (BASELINE configs[1]: 4096 channels x 65 536 samples = 128 warps for 148 SMs): a LINEAR or
 * AFFINE tick (zg_graph_linearity) is cut in time as well, one warp per (32 channels, time segment), in ZG_MODE_FAST --
 * a time-parallel evaluation re-associates the arithmetic, so ZG_MODE_EXACT always stays serial per channel
 * (flowz.hpp:1031-1074 evaluates the recurrence sample after sample; state-space reading: experimental_steps/
 * tuprix.cpp:239-254).  Two forms:
 *   ZG_TP_WARMUP    one launch.  Segment g starts K samples early from zero state and discards those outputs; K is
 *                   derived from the parameter values (float64, every channel): the smallest multiple of 4 boxes with
 *                   |A^K|_inf <= 2^-30, A = the tick's state matrix -- what is left of the true state after K ticks is
 *                   below fp32 resolution.  8 + 4K/L bytes per sample.  Not in place; ZG_ERR_UNSUPPORTED if some
 *                   channel does not forget within 8192 samples (poles on the unit circle: oscillators).
 *   ZG_TP_TWO_PASS  any linear tick: pass 1 runs every segment from zero state and keeps its final state, a small
 *                   kernel applies x <- A^L x + z along the boundaries, pass 2 runs every segment from its true
 *                   state.  12 bytes per sample (the input is read twice).
 *   ZG_TP_AUTO      warm-up form when there are too few channels to fill the GPU with one lane each, the block is long
 *                   enough (segments >= 8 K) and K exists; else, for graphs without a section-parallel kernel, the
 *                   two-pass form; else serial.          ZG_TP_OFF: never cut time.                              */
typedef enum zg_time_parallel { ZG_TP_AUTO = 0, ZG_TP_OFF = 1, ZG_TP_WARMUP = 2, ZG_TP_TWO_PASS = 3 } zg_time_parallel;

#define ZG_MAX_WIRES 8

typedef struct zg_plan_opts {
    int device;           /* CUDA device ordinal                                                */
    int64_t channels;     /* C: independent voices evaluated by this plan                       */
    int mode;             /* zg_mode                                                            */
    int layout;           /* zg_layout                                                          */
    int io_dtype;         /* sample storage in HBM: ZG_F32, or ZG_BF16 (state, parameters and all
                             arithmetic stay fp32; outputs are rounded to nearest even; generated
                             kernel only: no K1b / FIR kernel)                                     */
    int lanes_per_channel;/* 0 = auto; 1 = one lane per channel; S = S lanes per channel, lane k evaluating
                             section k of an S-section biquad cascade as a systolic pipeline (same
                             arithmetic, bit-identical in EXACT mode; planar layout, S = 2 or 4).  Auto
                             picks S when there are too few channels to fill the GPU with one lane each */
    int input_kind[ZG_MAX_WIRES];
    int force_jit;        /* 1 = never use the prebuilt biquad kernels (tests)                  */
    int time_parallel;    /* zg_time_parallel (ZG_MODE_FAST only; ignored -- serial -- in ZG_MODE_EXACT when AUTO) */
    int fir_tensor_cores; /* dense FIR in ZG_MODE_FAST: 0 = auto (planar fp32, 2..256 taps, blocks >= 256 samples run the
                             banded-Toeplitz contraction on tcgen05 tensor cores, 3xTF32: ~3e-6 block-relative from the
                             reference's left-to-right sum, dominated by the tensor core's truncating fp32 accumulation),
                             1 = never (CUDA-core FMA kernel, ~3e-7).  ZG_MODE_EXACT never uses tensor cores.        */
    int section_warps;    /* biquad cascades, planar fp32, blocks of whole 32-sample boxes: the sections of a group of 32
                             channels spread over the warps of a persistent CTA that share one ring of tiles (same
                             arithmetic, bit-identical in EXACT mode), every channel row moved in runs of 1 KB and more,
                             rows handed from one CTA to the next in the middle of a block.  0 = auto (4 sections: every
                             channel count; 3, 5-8 sections: from ~9500 channels on; blocks of at least 8 tiles),
                             1 = never, 2 = whenever the shape allows.  Calls on one plan must be stream-ordered (as for
                             any plan: its state rows live in device memory).                                     */
    int reserved[4];
} zg_plan_opts;

void zg_plan_opts_default(zg_plan_opts* o);
/* The kernel zg_plan_create() would specialise for this graph (the generated-tick path), as CUDA
 * source (want_cubin = 0) or as an sm_100a cubin (want_cubin = 1).  Needs libnvrtc only -- no
 * device -- so it doubles as an offline build check.  buf may be NULL to query *size.           */
int zg_graph_kernel_compile(const zg_graph* g, const zg_plan_opts* opts, int uniform_params, int want_cubin,
                            char* buf, size_t capacity, size_t* size);
/* Environment: ZG_KERNEL_CACHE_DIR=<dir> keeps the NVRTC-built cubins of generated kernels across processes
 * (keyed by the specialised source, the contraction mode and the NVRTC version); without it they are cached for
 * the life of the process only.                                                                   */
int zg_plan_create(const zg_graph* g, const zg_plan_opts* opts, zg_plan** out);
void zg_plan_destroy(zg_plan* p);

typedef struct zg_plan_info {
    char kernel[96];      /* name of the kernel this plan launches                              */
    int jit;              /* 1 = specialised with NVRTC at plan time, 0 = prebuilt in the library */
    int lanes_per_channel;/* lanes that evaluate one channel (1, or the section count: see zg_plan_opts) */
    int host_chunks;      /* row chunks the last zg_process_host call streamed its block in         */
    int regs_per_thread;
    int smem_bytes;
    int launches;         /* kernels launched by this plan so far                               */
    int threads_per_cta;  /* geometry of the last launch                                        */
    int stages;           /* TMA pipeline depth per warp of the last launch (FIR kernel: boxes in
                             the CTA's input ring)                                              */
    int uniform_params;   /* 1 = all parameters are scalars and travel in the constant bank     */
    int boxes;            /* 4 KB boxes (32 channels x 32 samples) per wire per pipeline stage
                             (FIR kernel: boxes per time segment)                               */
    int time_segments;    /* segments the last launch cut the block into (1 = serial in time)    */
    int segment_samples;  /* L: samples per segment                                              */
    int warmup_samples;   /* K of the warm-up form (0: serial or two-pass)                       */
    int linearity;        /* zg_linearity of the tick                                            */
} zg_plan_info;
int zg_plan_get_info(const zg_plan* p, zg_plan_info* info);

/* One block of n_samples for all channels.  in[i] / out[j] are DEVICE pointers to sample buffers
 * (io_dtype) in the plan's layout; ld is the leading dimension in elements (planar: >= n_samples,
 * interleaved: >= channels; a multiple of 16 bytes).  Buffers must be 16-byte aligned.  Asynchronous on
 * `stream` (a cudaStream_t, may be NULL).  Consecutive calls continue the stream exactly like
 * consecutive ticks: state is read at block start and written back at block end.
 * In place: out[j] may be the same buffer as in[i] (same pointer and ld; any pairing, e.g. two wires swapped) --
 * the block is then overwritten with the result, like `x[t] = f(x[t])` around the reference's tick.  Buffers that
 * overlap in any other way are ZG_ERR_ARG; the FIR kernel does not run in place (ZG_ERR_UNSUPPORTED).          */
int zg_process(zg_plan* p, const void* const* in, void* const* out, int64_t n_samples,
               int64_t ld_in, int64_t ld_out, void* stream);
/* Same, with HOST pointers: copies in, runs, copies out, synchronises (what a CPU-side caller of
 * the reference would switch to).                                                              */
int zg_process_host(zg_plan* p, const void* const* in, void* const* out, int64_t n_samples,
                    int64_t ld_in, int64_t ld_out);

/* state: [n_state][channels] floats, slot order as in zg_graph_dump(); host pointers */
int zg_state_reset(zg_plan* p);
int zg_state_get(zg_plan* p, float* host, size_t n_floats);
int zg_state_set(zg_plan* p, const float* host, size_t n_floats);
/* $index for all channels: n == 1 broadcasts a scalar, n == channels sets one value per channel */
int zg_param_set(zg_plan* p, int index, const float* host_values, int64_t n);
/* Same from DEVICE memory, one value per channel (n == channels): coefficients computed on the GPU -- e.g. RBJ
 * sections from per-voice f, Q (reactive_equations/reactive_filter_coeff.cpp:16-50) -- never visit the host.
 * Synchronises the device (the values are copied before the call returns).                               */
int zg_param_set_device(zg_plan* p, int index, const float* device_values, int64_t n);

/* ---- channel sharding across the GPUs of a box --------------------------------------------------
 * Voices are independent (each is its own stateful_lambda, flowz.hpp:1181-1230), so rank r of world_size owns the
 * contiguous channel range [*begin, *end) with its own plan; sizes differ by at most one, the first ranks are the
 * larger ones.  The library itself is single-device: the exchange at the edges (one scatter of the input block, one
 * gather of the output block, or none when the kernels work on the root's blocks through peer mappings) belongs to
 * the caller's process group -- zignal_b200/shard.py does it over torch.distributed.                             */
int zg_shard_range(int64_t channels, int world_size, int rank, int64_t* begin, int64_t* end);

/* ---- how a launch of the section-split biquad kernel (zg_plan_opts.section_warps) would be laid out ----------------
 * Host-only (needs no device): the geometry the launch planner picks for a planar fp32 cascade of `sections` direct-
 * form-1 biquads over `channels` x `samples` on a GPU of `sm_count` SMs with `max_smem` bytes of shared memory per
 * CTA; `segments` > 1 asks for the block cut in time (FAST, warm-up of `warmup_samples`).  Returns 1 and fills *out
 * when that kernel would run, 0 when the block stays on the other biquad kernels.  What the persistent CTAs rely on
 * -- at least as many rows of 32 channels (x segments) as groups of warps, so that a row is handed from one group to
 * the next at most once; a ring that fits; runs of boxes that divide the tile -- is checked over thousands of shapes
 * without a GPU (tests/test_split_plan.py).                                                                        */
typedef struct zg_split_plan {
    int groups_per_cta, warps_per_group, sections_per_warp, grid, threads_per_cta;
    int stages, boxes_per_tile, boxes_per_handover, smem_bytes;
    int segments, segment_boxes, warmup_boxes;
} zg_split_plan;
int zg_split_plan_query(int sections, int exact, int64_t channels, int64_t samples, int sm_count, int max_smem,
                        int segments, int warmup_samples, zg_split_plan* out);

#ifdef __cplusplus
}
#endif
#endif /* ZIGNAL_B200_H */
