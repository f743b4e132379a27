/*
 * tegb200.h -- C ABI of libtegb200.so, the thin runtime behind Teg's CUDA evaluation backend.
 *
 * This is the drop-in boundary for the hot path "evaluate a reduced Teg program over a batch of
 * independent instances".  Plain C types only (pointers, sizes, doubles); no torch / C++ types.
 * Every entry point returns 0 on success and a non-zero tegb_status otherwise; the message is then
 * available from tegb_last_error() (thread-local).  Nothing here ever abort()s.
 *
 * What each entry point replaces in the reference (/root/reference):
 *
 *   tegb_compile            C_EvalMode._preprocess + _compile: write tegout_N.h / main_N.cc and shell out
 *                           to `g++ -O3 -std=c++11 -fPIC` (teg/eval/c_eval.py:100-127).  Here: NVRTC,
 *                           sm_100a, --fmad=false, in process, result is a cubin.
 *   tegb_program_create     the `main()` the reference generates around the entry function
 *                           (teg/eval/c_eval.py:58-98): binds the entry points and argument layout.
 *   tegb_program_launch /   C_EvalMode.eval (teg/eval/c_eval.py:129-157): one Popen of the compiled exe
 *   tegb_program_eval_host  per instance with argv marshalling and stdout parsing.  Here: ONE launch for
 *                           n instances, structure-of-arrays inputs/outputs.
 *   tegb_render_launch      the per-pixel Python double loop of tests/plot.py:8-29 (render_image):
 *                           pixel boxes are generated on the device from the instance index.
 *   tegb_reduce_sum         (new) per-parameter sum of reverse-mode gradients over instances.
 *   tegb_allreduce_sum      (new) the one collective of the multi-GPU path (NCCL sum of the [P] vector).
 *   tegb_last_error         replaces `assert err is None` (c_eval.py:125,147), which can never fire.
 *
 * Layout: argument a of instance i is in_ptrs[j][i] (j = position of a among the array arguments);
 * output component k of instance i is out_ptrs[k][i].  Element type is float or double per program.
 */
#ifndef TEGB200_H
#define TEGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TEGB_OK = 0,
    TEGB_ERR_INVALID = 1,      /* bad argument */
    TEGB_ERR_COMPILE = 2,      /* NVRTC rejected the source; tegb_last_error() holds the log */
    TEGB_ERR_CUDA = 3,         /* driver / runtime error */
    TEGB_ERR_NO_DEVICE = 4,    /* no usable CUDA device or driver */
    TEGB_ERR_NCCL = 5,
    TEGB_ERR_PEER = 6          /* a rank of a peer group did not arrive within the time-out: sums were poisoned (NaN) */
} tegb_status;

typedef struct tegb_module tegb_module;     /* one compiled translation unit (cubin) */
typedef struct tegb_program tegb_program;   /* uniform + main kernel of one Teg program, loaded on a device */

/* how the runtime obtains an argument that is not a per-instance array */
typedef struct {
    const char *uniform_kernel;   /* <name>_uniform */
    const char *main_kernel;      /* <name>_main */
    const char *uniform_symbol;   /* name of the __constant__ table ("tegU") */
    int elem_size;                /* 4 = float, 8 = double */
    int n_uniform;                /* entries of the uniform table */
    int n_scalar_args;            /* by-value arguments of the uniform kernel */
    int n_array_args;             /* pointer arguments of the main kernel */
    int n_grid_args;              /* arguments generated from the pixel index (0 unless a render program) */
    int n_outputs;                /* output streams of the main kernel */
    int block_size;               /* threads per block the main kernel was generated for */
    int reduce_outputs;           /* number R of TRAILING outputs that are summed over the instances (0 = none,
                                     n_outputs = all): the main kernel writes per-block partial sums for them and
                                     launches deliver the R sums to out_ptrs[n_outputs - R] (one device pointer);
                                     out_ptrs[0 .. n_outputs - R) stay ordinary per-instance output streams */
    int team;                     /* threads cooperating on one instance (power of two, 1 = thread per instance);
                                     > 32 means one block of `team` threads per instance (block_size == team) */
    int vec;                      /* instances per thread of a streaming program (1 or 4); > 1 requires every
                                     input / output pointer to be 16-byte aligned */
    int grouped;                  /* 1: grouped render program (tegb_render_groups_launch): the scalar arguments are
                                     per-group arrays, the uniform table has one row per group */
    int n_pixel_args;             /* grouped programs: arguments read from images indexed by pixel */
    int accumulate;               /* grouped programs: outputs are added into the images in out_ptrs */
    int n_tile_conds;             /* render programs: tile conditions classified by `tile_kernel` before every
                                     launch (0 = none).  A tile is tile_cols pixel columns x tile_rows pixel rows;
                                     the kernel writes one tri-state (0 false / 1 true for every pixel of the tile,
                                     2 unresolved) per condition and tile, which the main kernel reads per block */
    int tile_rows;                /* render programs: pixel rows walked by one block (>= 1) */
    int tile_cols;                /* tile width: a multiple of 32 that divides block_size (whole warps per tile) */
    int tile_samples;             /* samples per integral the unit was generated for (the classification bounds the
                                     midpoint abscissae lb + step * (i + 0.5), i < tile_samples, of every pixel) */
    const char *tile_kernel;      /* <name>_tiles, or NULL when n_tile_conds == 0 */
} tegb_program_desc;

/* pixel grid of tests/plot.py:11-16: np.linspace(lo, hi, res+1) per axis, instance = nx * res_y + ny */
typedef struct {
    double x_lo, x_hi, y_lo, y_hi;
    int32_t res_x, res_y;
    int32_t row_begin, row_end;   /* rows nx in [row_begin, row_end) are evaluated (multi-GPU sharding) */
} tegb_grid;

const char *tegb_last_error(void);
int tegb_version(void);
/* hash of the runtime header (teg_b200_runtime.cuh) embedded in the library: generated units include it by name, so it
 * belongs in any cache key of compiled units */
uint64_t tegb_runtime_header_hash(void);
/* version of the NVRTC the library loaded, major * 1000 + minor * 10 (e.g. 12090); 0 if none could be loaded */
int tegb_nvrtc_version(void);

/* number of CUDA devices visible (0 without a driver); never fails */
int tegb_device_count(void);

/* --- JIT ---------------------------------------------------------------------------------- */
/* CUDA text -> cubin for `arch` ("sm_100a").  Needs no GPU.  `opts`: extra NVRTC options separated by
 * spaces (may be NULL).  The runtime header teg_b200_runtime.cuh is supplied by the library. */
int tegb_compile(const char *cuda_src, const char *arch, const char *opts, tegb_module **out);
int tegb_module_from_cubin(const void *cubin, size_t nbytes, tegb_module **out);
int tegb_module_cubin(const tegb_module *m, const void **cubin, size_t *nbytes);
const char *tegb_module_log(const tegb_module *m);
void tegb_module_destroy(tegb_module *m);

/* --- programs ----------------------------------------------------------------------------- */
/* Threads and streams: handles are independent of one another (each owns its copy of the module, its published
 * uniform table and its scratch buffers), so one handle per thread / stream runs fully concurrently.  ONE handle may
 * also be shared: its launch calls are serialised by a mutex, and a launch on another stream than the handle's
 * previous launch first waits, on the device, for that previous launch (the scratch is single-buffered).  The
 * reference's executor is "NOT thread-safe" (teg/eval/c_eval.py:55-56). */
int tegb_program_create(tegb_module *m, const tegb_program_desc *desc, int device, tegb_program **out);
void tegb_program_destroy(tegb_program *p);

/* Device buffers.  scalars[n_scalar_args] are converted to the program's element type.
 * in_ptrs[n_array_args], out_ptrs[n_outputs]: device pointers to n elements each.  Asynchronous on
 * `stream` (a cudaStream_t; NULL = default stream).  With reduce_outputs = R > 0 the last R outputs are
 * delivered as sums over the n instances to out_ptrs[n_outputs - R] (a device pointer to R elements; fixed
 * summation order for a given n: deterministic). */
int tegb_program_launch(tegb_program *p, const double *scalars, const void *const *in_ptrs,
                        void *const *out_ptrs, int64_t n, void *stream);

/* Host buffers: copies inputs H2D, launches, copies outputs D2H, synchronises.  Device staging buffers
 * are owned by the program and grow on demand. */
int tegb_program_eval_host(tegb_program *p, const double *scalars, const void *const *host_in,
                           void *const *host_out, int64_t n);

/* Render: the program's grid arguments (pixel box bounds) are generated in-kernel from the instance
 * index; out_ptrs[k] receive (row_end-row_begin)*res_y elements. */
int tegb_render_launch(tegb_program *p, const double *scalars, const tegb_grid *grid,
                       void *const *out_ptrs, void *stream);
/* the same, restricted to the tile rows tile_phase, tile_phase + tile_stride, ... of the row range (a tile row =
 * tile_rows consecutive pixel rows): interleaved sharding over ranks (tile_stride = world size, tile_phase = rank), which
 * gives every rank the same mix of cheap and expensive tiles.  The output images hold the rendered rows only, compact
 * and in order. */
int tegb_render_launch_strided(tegb_program *p, const double *scalars, const tegb_grid *grid, int tile_stride,
                               int tile_phase, void *const *out_ptrs, void *stream);
/* diagnostics: copies the tri-states written by the last tegb_render_launch of a program with tile conditions
 * (condition-major, [n_tile_conds][tiles], element type of the program) to host memory */
int tegb_program_tile_table(tegb_program *p, void *host_out, size_t nbytes, void *stream);

/* --- all-reduce over NVLink peer memory (one node) --------------------------------------------------------------
 * Replaces "column sums, then ncclAllReduce" for the per-parameter gradients of a sharded render:
 * every rank owns a mailbox mapped by all peers (CUDA IPC); the last column-sum kernel writes its sums into every
 * mailbox and adds the ranks' contributions in rank order (bit-identical on every rank).
 *   tegb_peer_create   allocates this rank's mailbox (up to max_values numbers per call) and returns its 64-byte handle
 *   tegb_peer_connect  maps the other ranks' mailboxes; `handles` = world x 64 bytes in rank order (the caller gathers
 *                      them, e.g. with torch.distributed.all_gather)
 *   tegb_program_set_peer  reduce_outputs launches of the program then deliver the all-reduced sums
 *   tegb_peer_allreduce_sum  stand-alone in-place all-reduce of a device vector (every rank must call it, same order)
 *   tegb_program_set_peer_groups  the same for grouped reduce programs: out_ptrs[0] of tegb_render_groups_launch then
 *                      receives the all-reduced [n_groups_total][n_outputs] sums; local_of (device int32
 *                      [n_groups_total] or NULL) maps every global group to this rank's group index (-1: not held)
 *   tegb_peer_error    1 if an exchange timed out (a peer never arrived); tegb_peer_destroy after a barrier
 * Time-out: an exchange waits TEGB200_PEER_TIMEOUT_S seconds (default 120; 0 = for ever) for every peer.  When it
 * expires the sums of that call are NaN, a sticky flag is raised and EVERY later launch on the group returns
 * TEGB_ERR_PEER (checked on the host before enqueueing, no synchronisation). */
typedef struct tegb_peer tegb_peer;
int tegb_peer_create(int rank, int world, int max_values, int device, tegb_peer **out, void *handle64);
int tegb_peer_connect(tegb_peer *p, const void *handles);
int tegb_peer_allreduce_sum(tegb_peer *p, void *buf, int count, int elem_size, void *stream);
int tegb_peer_error(tegb_peer *p);
void tegb_peer_destroy(tegb_peer *p);
int tegb_program_set_peer(tegb_program *prog, tegb_peer *peer);
int tegb_program_set_peer_groups(tegb_program *prog, tegb_peer *peer, int n_groups_total, const int32_t *local_of);

/* Grouped render: groups [group_begin, group_begin + group_count) (e.g. triangles) are evaluated over their
 * own pixel windows in ONE launch.  group_ptrs[n_scalar_args]: device arrays with one value per group
 * (indexed by absolute group number).  windows: device int32 [n_groups][4] = (row_begin, row_end, col_begin,
 * col_end) in pixels of the full grid; rows are clipped to the grid's [row_begin, row_end) band.
 * max_rows / max_cols: upper bounds of any window's height / width (launch geometry).  pixel_ptrs
 * [n_pixel_args]: device images of the band (row-major, (row - grid.row_begin) * res_y + col).
 * Outputs: images of the band (assigned, or added to when the program accumulates -- then the windows of
 * one launch must not overlap), or for reduce_outputs programs out_ptrs[0] -> [group_count][n_outputs]
 * sums over each group's window (row j of the result = group group_begin + j). */
int tegb_render_groups_launch(tegb_program *p, const void *const *group_ptrs, int group_begin, int group_count,
                              const tegb_grid *grid, const int32_t *windows, int max_rows, int max_cols,
                              const void *const *pixel_ptrs, void *const *out_ptrs, void *stream);

/* The per-group prologue of a grouped render (uniform rows + tile classification for the groups' CURRENT parameters)
 * as a call of its own: prepare a group range once, then launch any number of sub-ranges of it with
 * tegb_render_groups_launch_prepared (same grid, windows pointer and window bounds) -- e.g. the window-disjoint classes
 * of a scene, one accumulating launch each.  The parameters must not change between the prepare and the launches. */
int tegb_render_groups_prepare(tegb_program *p, const void *const *group_ptrs, int group_begin, int group_count,
                               const tegb_grid *grid, const int32_t *windows, int max_rows, int max_cols, void *stream);
int tegb_render_groups_launch_prepared(tegb_program *p, const void *const *group_ptrs, int group_begin, int group_count,
                                       const tegb_grid *grid, const int32_t *windows, int max_rows, int max_cols,
                                       const void *const *pixel_ptrs, void *const *out_ptrs, void *stream);

/* Groups with OVERLAPPING windows in one launch (accumulating image programs; the reference adds the triangles of a
 * scene one expression at a time, tests/plot.py:8-29 per triangle).  The launch covers the band in blocks of block_size
 * columns x tile_rows rows; block (bx, by) -- bin by * ceil(res_y / block_size) + bx -- walks the groups
 * bin_groups[bin_begin[bin] .. bin_begin[bin + 1]) (device int32 arrays, built by the caller from the windows: every group
 * whose window reaches into the block, ASCENDING), so every pixel is owned by one thread for the whole launch and adds
 * its contributions in ascending group order: the result of launching the groups one after the other, in one launch and
 * without the read-modify-write passes.  clear != 0: the images are zeroed by the launch itself (every block clears its
 * pixels before it adds to them) instead of being added to.  Runs its own prologue (groups [0, group_count)). */
int tegb_render_groups_gather(tegb_program *p, const void *const *group_ptrs, int group_count, const tegb_grid *grid,
                              const int32_t *windows, int max_rows, int max_cols, const int32_t *bin_begin,
                              const int32_t *bin_groups, int clear, const void *const *pixel_ptrs, void *const *out_ptrs,
                              void *stream);

/* kernels launched by this library since load (for accounting) */
int64_t tegb_launch_count(void);

/* --- reductions / collectives ------------------------------------------------------------- */
/* out[k] = sum_i in_ptrs[k][i], k < n_cols: deterministic (fixed order for a given n), two launches.
 * `out` is a device pointer to n_cols elements.  workspace: device, >= tegb_reduce_workspace_bytes. */
size_t tegb_reduce_workspace_bytes(int n_cols, int64_t n, int elem_size);
int tegb_reduce_sum(const void *const *in_ptrs, int n_cols, int64_t n, int elem_size, void *out,
                    void *workspace, size_t workspace_bytes, void *stream);

/* NCCL: 128-byte unique id for bootstrap, communicator creation, in-place sum all-reduce. */
int tegb_nccl_unique_id(void *id128);
int tegb_nccl_comm_create(const void *id128, int world_size, int rank, int device, void **comm);
int tegb_nccl_comm_destroy(void *comm);
int tegb_allreduce_sum(void *buf, int64_t count, int elem_size, void *comm, void *stream);

/* --- measurement -------------------------------------------------------------------------- */
/* CUDA events around the MAIN kernel of every launch of this handle (not its prologue / classification / column
 * sums): tegb_program_main_kernel_ms waits for the last launch and returns that kernel's duration.  For rooflines. */
int tegb_program_enable_timing(tegb_program *p, int on);
int tegb_program_main_kernel_ms(tegb_program *p, double *ms);
/* dependent-chain FMA micro-benchmark: achieved TFLOP/s (2 flops per FMA) on `device` */
int tegb_fma_peak(int elem_size, int device, double *tflops, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* TEGB200_H */
