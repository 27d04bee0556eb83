/*
 * madgpu.h -- C-ABI of libmadgpu.so, the B200 (sm_100a) multigrid solver for the implicit
 * anisotropic-diffusion step of nellogrb/MultigridAnisotropicDiffusion.
 *
 * The reference has no FFI: its boundary is the C++ template API of
 * itk::MultigridAnisotropicDiffusionImageFilter.  This header is what a re-written
 * GenerateData() (include/itkMultigridAnisotropicDiffusionImageFilter.hxx in this repo) binds
 * to; every entry point cites the reference routine it replaces (paths relative to
 * /root/reference/include).
 *
 * Conventions: plain pointers and sizes only; return 0 on success, a negative MADGPU_E* code on
 * failure (never throws); madgpu_last_error() returns a human-readable message.  The caller owns
 * all host pointers; the library owns all device memory, streams and events.  One context is
 * used by one host thread at a time.  Images are x-fastest (ITK index[0] contiguous), tensors
 * are the ITK SymmetricSecondRankTensor buffer as is: dim*(dim+1)/2 scalars per voxel in the
 * order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)  [2-D: (0,0),(0,1),(1,1)].
 *
 * There is no CPU fallback: every call fails with MADGPU_ECUDA when no sm_100 device is usable.
 */
#ifndef MADGPU_H
#define MADGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MADGPU_VERSION 1

/* error codes */
#define MADGPU_OK 0
#define MADGPU_EINVAL (-1)   /* bad argument / unsupported size                      */
#define MADGPU_ECUDA (-2)    /* CUDA runtime error (message in madgpu_last_error)   */
#define MADGPU_ENOMEM (-3)   /* device or host allocation failed                     */
#define MADGPU_ESTATE (-4)   /* call sequence error (e.g. solve before set_tensor)  */
#define MADGPU_ESINGULAR (-5)/* coarsest-grid operator is singular                   */
#define MADGPU_ENUMERIC (-6) /* the relative residual became non-finite (diverged / overflowed); no image returned */

/* TSmootherType template argument of the reference filter
 * (itkMultigridAnisotropicDiffusionImageFilter.h:89-92) as a run-time tag. */
#define MADGPU_SMOOTHER_GS 0 /* mad::MultigridGaussSeidelSmoother -> multicolour Gauss-Seidel */
#define MADGPU_SMOOTHER_WJ 1 /* mad::MultigridWeightedJacobiSmoother                         */

/* enum CycleType { VCYCLE, FMG, SMOOTHER } (itkMultigridAnisotropicDiffusionImageFilter.h:123) */
#define MADGPU_CYCLE_V 0
#define MADGPU_CYCLE_FMG 1
#define MADGPU_CYCLE_SMOOTHER 2

#define MADGPU_MAX_LEVELS 32
#define MADGPU_MAX_STEPS 64

typedef struct madgpu_ctx madgpu_ctx;

/* Parameters = the filter's setters (itkMultigridAnisotropicDiffusionImageFilter.h:133-156) plus
 * the image geometry GenerateData() reads from its input (…Filter.hxx:131). */
typedef struct madgpu_params {
  int32_t struct_size;        /* sizeof(madgpu_params), ABI check                                  */
  int32_t dim;                /* 2 or 3                                                            */
  int32_t size[3];            /* voxels per axis, x fastest; size[2] ignored in 2-D               */
  double spacing[3];          /* image spacing (the tensor image's spacing is ignored, as in the reference) */
  double time_step;           /* SetTimeStep, default 0.01 (…Filter.hxx:39)                        */
  int32_t number_of_steps;    /* SetNumberOfSteps, default 1                                       */
  int32_t cycle;              /* SetCycle, default MADGPU_CYCLE_V                                  */
  int32_t iterations_per_grid;/* SetIterationsPerGrid, default 2                                   */
  double tolerance;           /* SetTolerance, default 1e-6                                        */
  int32_t max_cycles;         /* SetMaxCycles, default 100                                         */
  int32_t verbose;            /* SetVerbose, default 0 (per-cycle lines on stdout)                 */
  int32_t smoother;           /* MADGPU_SMOOTHER_*, default GS (the filter's default template arg) */
  double omega;               /* Jacobi weight, default 2/3 (mad/itkMultigridWeightedJacobiSmoother.hxx:186-191) */
  int32_t gs_colors;          /* 4 (default) or 8 colours for the 3-D multicolour sweep; 2-D always 4 */
  int32_t device;             /* CUDA device ordinal, default 0                                    */
  /* z-slab decomposition (one context per rank/GPU).  world_size == 1: whole volume.            */
  int32_t rank;
  int32_t world_size;
  int32_t reserved[8];
} madgpu_params;

/* Per-solve statistics (replaces the reference's BENCHMARK trace, …Filter.hxx:147-151, 401-409). */
typedef struct madgpu_stats {
  int32_t struct_size;
  int32_t steps;                              /* time steps executed                               */
  int32_t cycles_per_step[MADGPU_MAX_STEPS];  /* V-cycles (or smoother iterations) per step        */
  double final_relres[MADGPU_MAX_STEPS];      /* ||f-Au||/||f|| when the step stopped              */
  int32_t total_cycles;
  int32_t levels;
  double setup_ms;                            /* last set_tensor: ingest + restriction + coarse factorisation */
  double h2d_ms, d2h_ms;                      /* host<->device copies of the last solve (wall)     */
  double solve_ms;                            /* device time of the cycle loop (CUDA events)       */
  double fmg_ms;                              /* device time of the FMG prologue(s)                */
  int64_t kernel_launches;                    /* kernels launched by the last solve                */
  /* profiling (madgpu_set_profiling(ctx,1)): device ms and launch counts by kernel class          */
  double prof_ms[16];
  int64_t prof_launches[16];
  int64_t graph_launches;                     /* CUDA-graph replays of the captured coarse part of a cycle (their kernels are counted in kernel_launches) */
} madgpu_stats;

/* kernel classes for prof_ms / prof_launches */
#define MADGPU_K_SMOOTH0 0     /* smoother sweeps on level 0                   */
#define MADGPU_K_SMOOTHC 1     /* smoother sweeps on levels >= 1               */
#define MADGPU_K_RESID0 2      /* level-0 fp64 residual + norm (stop test)     */
#define MADGPU_K_RESTRICT 3    /* residual + restriction                       */
#define MADGPU_K_PROLONG 4     /* prolongation + correction                    */
#define MADGPU_K_COARSE 5      /* coarsest-grid solve                          */
#define MADGPU_K_MISC 6        /* fills, casts, axpy                           */
#define MADGPU_K_HALO 7        /* halo exchange (multi-GPU)                    */
#define MADGPU_K_GRAPH 8       /* captured coarse part of a V-cycle (CUDA graph): every class of the levels <= 128^3 voxels */

void madgpu_params_default(madgpu_params *p);

/* Replaces `new GridsHierarchyType(...)` minus the tensor (…Filter.hxx:131): level schedule
 * (mad/itkGridsHierarchy.hxx:36-106), device buffers for every level, stream. */
int madgpu_create(const madgpu_params *p, madgpu_ctx **out);
void madgpu_destroy(madgpu_ctx *ctx);
const char *madgpu_last_error(const madgpu_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* ---- z-slab decomposition over the GPUs of one node (new; the reference is single-process) ----
 * One process (and one context) per GPU.  params.size is the GLOBAL volume, params.rank / world_size say which slab of
 * z planes this context owns: size[2] / world_size consecutive planes, rank 0 lowest.  The finest levels stay
 * distributed (one xy-plane halo per neighbour and sweep over NCCL send/recv on the solver's stream, one scalar
 * all-reduce per norm); levels of 64^3 voxels or fewer are gathered onto rank 0, which runs the rest of the V-cycle
 * and scatters the correction back.  Every entry point becomes a collective: all ranks call it with their own slab
 * (images, tensor) in the same order.  Requirements: 3-D, size[2] divisible by world_size, planes per rank even
 * and >= 4.  NCCL is bound at run time (dlopen of libnccl.so.2).
 *   id128: 128 bytes from madgpu_nccl_unique_id() on one rank, distributed by the caller (MPI, torch.distributed, ...). */
int madgpu_nccl_unique_id(void *id128);
int madgpu_create_slab(const madgpu_params *p, const void *nccl_unique_id, madgpu_ctx **out);
/* Optional, after madgpu_create_slab on every rank: peer-memory halo.  Each rank exports the CUDA IPC handles of its
 * level fields (madgpu_ipc_export: pass blob = NULL to learn the size), the caller moves the blobs between the ranks,
 * and every rank imports the blobs of rank-1 and rank+1 (NULL where there is none).  From then on the kernel that
 * produces a field stores its two boundary planes straight into the neighbours' ghost planes over NVLink and the stream
 * bumps an arrival counter in the neighbours' memory (cuStreamWriteValue32 / cuStreamWaitValue32): no separate exchange
 * step.  NCCL send/recv remains the fallback (tensor set-up, agglomeration level, contexts that do not import). */
int madgpu_ipc_export(madgpu_ctx *ctx, void *blob, size_t capacity, size_t *needed);
int madgpu_ipc_import(madgpu_ctx *ctx, const void *blob_lower, const void *blob_upper);
/* import ends with a handshake with both neighbours and fails (MADGPU_ECUDA) when it does not complete; the ranks must then
 * agree: if any rank failed, all call madgpu_ipc_disable and the NCCL exchange stays in use.
 * Environment: MADGPU_P2P_WAIT=kernel awaits the arrival counters with a bounded one-thread kernel (MADGPU_P2P_TIMEOUT_MS, default
 * 10000) instead of cuStreamWaitValue32: a signal that never arrives then makes the running solve / cycles call fail with
 * MADGPU_ECUDA on every rank in the same cycle (the context must be recreated) instead of hanging the stream. */
int madgpu_ipc_disable(madgpu_ctx *ctx);
/* planes [z_begin, z_begin + z_count) of `level` held by this context; global_nz = planes of the whole level.
 * Valid for the levels this context holds (all of them when world_size == 1). */
int madgpu_slab(const madgpu_ctx *ctx, int32_t level, int32_t *z_begin, int32_t *z_count, int32_t *global_nz);

/* Run-time changes of the solver settings that do not alter the hierarchy
 * (itkSetMacro setters; time_step / size / spacing changes need a new context). */
int madgpu_set_solver(madgpu_ctx *ctx, int32_t smoother, double omega, int32_t iterations_per_grid, int32_t cycle,
                      double tolerance, int32_t max_cycles, int32_t number_of_steps, int32_t verbose);

/* Replaces SetDiffusionTensor (…Filter.hxx:66-101) + the tensor part of the GridsHierarchy constructor
 * (mad/itkGridsHierarchy.hxx:112-201: component split, per-level full-weighting restriction) +
 * DirectSolver's factorisation (mad/itkDirectSolver.hxx:32-88).  `aos` is a HOST pointer to the ITK
 * tensor buffer.  The operator rows (GenerateDCA, mad/itkGridsHierarchy.hxx:298-516) are never
 * materialised: kernels evaluate them from the six (three) tensor planes. */
int madgpu_set_tensor_f32(madgpu_ctx *ctx, const float *aos);
int madgpu_set_tensor_f64(madgpu_ctx *ctx, const double *aos);
/* Same, tensor already resident on the device as SoA fp32 planes (ncomp pointers, each
 * size[0]*size[1]*size[2] dense floats, x fastest). */
int madgpu_set_tensor_device_f32(madgpu_ctx *ctx, const float *const *planes);

/* Replaces GenerateData() (…Filter.hxx:104-297): input cast, time-step loop, V-cycle / FMG /
 * smoother-only iterations with the relative-residual stop test, output static_cast.
 * `in` and `out` are HOST pointers (may alias).  The suffix names the input pixel type; the output
 * pixel type is the same (the reference tests use TInputImage == TOutputImage), except
 * madgpu_solve_cast which takes both. */
#define MADGPU_PIX_U8 0
#define MADGPU_PIX_I16 1
#define MADGPU_PIX_F32 2
#define MADGPU_PIX_F64 3
int madgpu_solve_cast(madgpu_ctx *ctx, int32_t in_type, const void *in, int32_t out_type, void *out, madgpu_stats *stats);
int madgpu_solve_u8(madgpu_ctx *ctx, const uint8_t *in, uint8_t *out, madgpu_stats *stats);
int madgpu_solve_i16(madgpu_ctx *ctx, const int16_t *in, int16_t *out, madgpu_stats *stats);
int madgpu_solve_f32(madgpu_ctx *ctx, const float *in, float *out, madgpu_stats *stats);
int madgpu_solve_f64(madgpu_ctx *ctx, const double *in, double *out, madgpu_stats *stats);
/* Same with DEVICE pointers to dense fp32 images (inputs already resident in HBM).  d_in == NULL continues from the fp64 result of
 * the previous solve of this context instead of an fp32 image: the outer iterations of VEDMultigridImageFilter::GenerateData carry
 * the image in double (itkVEDMultigridImageFilter.h:64, .hxx:105-123). */
int madgpu_solve_device_f32(madgpu_ctx *ctx, const float *d_in, float *d_out, madgpu_stats *stats);

/* Cycle-level driving of the same loop (benchmarks, per-V-cycle parity): begin stages the image and
 * forms the first residual (…Filter.hxx:182-204), run executes exactly n outer iterations -- V-cycle
 * (or smoother sweep in SMOOTHER mode) + fp64 residual + norm read-back, i.e. one pass of the do-while
 * body (…Filter.hxx:207-246) each -- ignoring tolerance, and reports the device time of the n
 * iterations measured with CUDA events on the library's stream; end casts the iterate out.
 * relres (n doubles) and device_ms may be NULL. */
int madgpu_cycles_begin_device_f32(madgpu_ctx *ctx, const float *d_in);
int madgpu_cycles_begin_f32(madgpu_ctx *ctx, const float *in);
int madgpu_cycles_run(madgpu_ctx *ctx, int32_t n, double *relres, float *device_ms, madgpu_stats *stats);
int madgpu_cycles_end_device_f32(madgpu_ctx *ctx, float *d_out);
int madgpu_cycles_end_f64(madgpu_ctx *ctx, double *out);

/* The current iterate (the result of the last solve / cycles_run) cast to a HOST buffer of pixel type out_type
 * (MADGPU_PIX_*): the output cast of GenerateData (…Filter.hxx:267-284) for callers that drove the solve with device
 * buffers (madgpu_solve_device_f32, madved_run) and want the fp64 iterate cast once, not via fp32. */
int madgpu_fetch_output(madgpu_ctx *ctx, int32_t out_type, void *out);

/* relative residual after every cycle of the last solve: hist[step * max_cycles + cycle] */
int madgpu_get_relres_history(const madgpu_ctx *ctx, double *hist, int32_t capacity);

/* on: bit mask of kernel classes (1 << MADGPU_K_*) to time with CUDA events; 0 = off, -1 = all */
int madgpu_set_profiling(madgpu_ctx *ctx, int32_t on);

/* ---- hierarchy introspection (GridsHierarchy getters, mad/itkGridsHierarchy.h:86-113) ---- */
int madgpu_num_levels(const madgpu_ctx *ctx);
/* size / spacing / centering (0 vertex, 1 cell: how level l was obtained from l-1) of level l */
int madgpu_level_info(const madgpu_ctx *ctx, int32_t level, int32_t size[3], double spacing[3], int32_t centering[3]);

/* Ordering of the Gauss-Seidel sweep on a level (the reference sweeps lexicographically,
 * mad/itkMultigridGaussSeidelSmoother.h:87-100).  tile = {0,0,0}: multicolour, one pass per colour over
 * the whole level.  Otherwise the fused sweep: the level is cut in tiles of tile[0] x tile[1] x tile[2]
 * voxels (x, y, z); inside a tile planes are relaxed in z order, each plane as even rows (even x, odd x)
 * then odd rows; values outside the tile are those of the previous sweep. */
int madgpu_gs_tile(const madgpu_ctx *ctx, int32_t level, int32_t tile[3]);

/* The passes the NEXT n_iter Gauss-Seidel sweeps of a leg on `level` will be run as (no reference counterpart: the reference runs
 * n_iter lexicographic sweeps, mad/itkMultigridGaussSeidelSmoother.hxx:33-111).  Six values per pass: {sweeps fused in the pass,
 * tile x, tile y, tile z, tile-grid shift y, shift z}.  A pass that fuses S > 1 sweeps (temporal blocking) runs S sweeps of the
 * ordering above inside every tile with the values outside the tile frozen at those the pass started from; the tile of voxel
 * (y, z) is ((y + shift y) / tile y, (z + shift z) / tile z).  Returns the number of passes, or a negative error. */
int madgpu_gs_leg_plan(const madgpu_ctx *ctx, int32_t level, int32_t n_iter, int32_t *passes, int32_t capacity);

/* ---- per-operator entry points (isolated parity tests; HOST dense fp32 buffers of the level's size) ---- */
/* restricted tensor planes of a level: ncomp * nvox floats, SoA (mad/itkGridsHierarchy.hxx:149-162) */
int madgpu_op_get_tensor(madgpu_ctx *ctx, int32_t level, float *planes);
/* explicit operator rows as the kernels evaluate them: nvox * 3^dim floats in Neighborhood raster
 * order (GenerateDCA, mad/itkGridsHierarchy.hxx:298-516) */
int madgpu_op_assemble(madgpu_ctx *ctx, int32_t level, float *stencil);
/* n_iter smoother iterations (SingleIteration: mad/itkMultigridWeightedJacobiSmoother.hxx:33-102,
 * mad/itkMultigridGaussSeidelSmoother.hxx:33-111 -> multicolour) */
int madgpu_op_smooth(madgpu_ctx *ctx, int32_t level, int32_t smoother, int32_t n_iter, const float *u, const float *f,
                     float *out);
/* r = f - A u (ComputeResidual, mad/itkMultigridGaussSeidelSmoother.hxx:114-180) and ||r||_2
 * (L2Norm, …Filter.hxx:496-515); r or norm may be NULL */
int madgpu_op_residual(madgpu_ctx *ctx, int32_t level, const float *u, const float *f, float *r, double *norm);
/* level-0 fp64 residual used by the stop test: u, f, r are HOST doubles */
int madgpu_op_residual_f64(madgpu_ctx *ctx, const double *u, const double *f, double *r, double *norm);
/* coarse(level+1) = R fine(level)  (Restriction, mad/itkInterGridOperators.hxx:175-304) */
int madgpu_op_restrict(madgpu_ctx *ctx, int32_t fine_level, const float *fine, float *coarse);
/* fine(level) = P coarse(level+1)  (Interpolation, mad/itkInterGridOperators.hxx:45-172) */
int madgpu_op_prolong(madgpu_ctx *ctx, int32_t fine_level, const float *coarse, float *fine);
/* e = A_L^-1 f on the coarsest level (DirectSolver::Solve, mad/itkDirectSolver.hxx:91-147) */
int madgpu_op_coarse_solve(madgpu_ctx *ctx, const float *f, float *e);
/* one V-cycle started at `level` (VCycle, …Filter.hxx:341-493) */
int madgpu_op_vcycle(madgpu_ctx *ctx, int32_t level, const float *u, const float *f, float *out);

#ifdef __cplusplus
}
#endif
#endif /* MADGPU_H */
