/*
 * madved.h -- C-ABI of the VED tensor front-end in libmadgpu.so (B200, sm_100a): the part of
 * itk::VEDMultigridImageFilter that runs immediately BEFORE the multigrid diffusion solve of madgpu.h --
 * Hessian at several scales, vesselness, diffusion-tensor synthesis.
 * With it the tensor is produced in HBM and handed to the solver there (madgpu_set_tensor_device_f32), so the
 * 6 x N x 8-byte tensor upload of the host path disappears and a whole VED filter run stays on the device.
 *
 * Every entry point cites the reference routine it replaces (paths relative to /root/reference/include).
 * Conventions as in madgpu.h: plain pointers and sizes, 0 / negative MADGPU_E* codes, never throws, no CPU
 * fallback (MADGPU_ECUDA without a usable sm_100 device).  Volumes are 3-D, x fastest; Hessians and tensors are
 * the ITK SymmetricSecondRankTensor buffer: six scalars per voxel, (0,0),(0,1),(0,2),(1,1),(1,2),(2,2).
 *
 * Third-party arithmetic: the reference computes the Hessian with itk::HessianRecursiveGaussianImageFilter and the
 * eigen-system with vnl_symmetric_eigensystem, neither of which is part of the reference's sources.  The kernels
 * restate the published algorithms (4th-order Deriche recursions with ITK's normalisation; a symmetric 3x3 eigen-solve);
 * PARITY OF THOSE TWO IS UNPINNED.  A caller that wants ITK's own Hessian keeps it and
 * passes the result to madved_update_vesselness_host_f64.
 */
#ifndef MADVED_H
#define MADVED_H

#include <stddef.h>
#include <stdint.h>

#include "madgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct madved_ctx madved_ctx;

/* The VED setters (itkVEDMultigridImageFilter.h:88-93) plus the image geometry. */
typedef struct madved_params {
  int32_t struct_size; /* sizeof(madved_params), ABI check                               */
  int32_t size[3];     /* voxels per axis, x fastest; every axis >= 4 (the recursive filter's minimum line length) */
  double spacing[3];
  double alpha;        /* SetAlpha, default 0.5        (itkVEDMultigridImageFilter.hxx:36) */
  double beta;         /* SetBeta, default 0.5         (:37)                              */
  double gamma;        /* SetGamma, default 5.0        (:38)                              */
  double epsilon;      /* SetEpsilon, default 0.01     (:39)                              */
  double omega;        /* SetOmega, default 5.0        (:40)                              */
  double sensitivity;  /* SetSensitivity, default 10.0 (:41)                              */
  int32_t device;      /* CUDA device ordinal                                             */
  int32_t reserved[7];
} madved_params;

typedef struct madved_stats {
  int32_t struct_size;
  int32_t scales;          /* Hessians consumed since madved_begin                          */
  double hessian_ms;       /* device time of the recursive-Gaussian passes since madved_begin */
  double vesselness_ms;    /* device time of the eigen / vesselness / tensor kernel         */
  double h2d_ms, d2h_ms;   /* host<->device copies of the last madved_run (wall)            */
  double diffusion_ms;     /* wall time of the DiffusionStep calls of the last madved_run   */
  int64_t kernel_launches; /* kernels launched by this context since madved_begin (madved_run: whole run, solver included) */
} madved_stats;

void madved_params_default(madved_params *p);
int madved_create(const madved_params *p, madved_ctx **out);
void madved_destroy(madved_ctx *ctx);
const char *madved_last_error(const madved_ctx *ctx); /* ctx may be NULL: error of the last failed create */
int madved_set_params(madved_ctx *ctx, double alpha, double beta, double gamma, double epsilon, double omega, double sensitivity);

/* The filter's internal image (GenerateData's cast of the input, itkVEDMultigridImageFilter.hxx:70-100).
 * type: MADGPU_PIX_*; HOST pointer, or a DEVICE pointer to dense fp32. */
int madved_set_image(madved_ctx *ctx, int32_t type, const void *image);
int madved_set_image_device_f32(madved_ctx *ctx, const float *d_image);
/* device pointer of that image (dense fp32): the in/out argument of madgpu_solve_device_f32 for DiffusionStep */
int madved_image_device(madved_ctx *ctx, float **d_image);

/* Forget the vesselness state: m_MaxVesselnessResponse = 0 etc. (:121-123); the next Hessian is "the first". */
int madved_begin(madved_ctx *ctx);

/* ComputeHessian (:158-173): Hessian of the current image at scale sigma (physical units), NormalizeAcrossScale on,
 * into the context's six Hessian planes. */
int madved_hessian(madved_ctx *ctx, double sigma);
/* UpdateVesselness (:215-299) on those planes: per voxel eigen-system, magnitude sort, VesselnessFunction (:176-212),
 * keep the best scale; the tensor of the best scale (GenerateDiffusionTensor, :302-378) is updated in the same pass. */
int madved_update_vesselness(madved_ctx *ctx);
/* Same, with a Hessian computed by the caller (e.g. ITK's own filter): HOST buffer, 6 doubles per voxel. */
int madved_update_vesselness_host_f64(madved_ctx *ctx, const double *hessian_aos);
/* madved_hessian + madved_update_vesselness */
int madved_add_scale(madved_ctx *ctx, double sigma);

/* Results.  planes[6]: DEVICE pointers to the dense fp32 tensor planes (xx,xy,xz,yy,yz,zz), the argument of
 * madgpu_set_tensor_device_f32; valid until the context is destroyed, contents change with every update. */
int madved_tensor_planes(madved_ctx *ctx, const float **planes);
int madved_get_tensor_f64(madved_ctx *ctx, double *tensor_aos);   /* HOST, 6 doubles per voxel  */
int madved_get_response_f64(madved_ctx *ctx, double *response);   /* HOST, m_MaxVesselnessResponse */
int madved_get_hessian_f64(madved_ctx *ctx, double *hessian_aos); /* HOST, the last madved_hessian */
int madved_get_stats(madved_ctx *ctx, madved_stats *stats);

/* VEDMultigridImageFilter::GenerateData (:63-155) entirely on the device: cast in, `iterations` times
 * { for every scale: Hessian + vesselness; tensor; DiffusionStep (:381-402) }, cast out (static_cast, :141).
 * `solver` is a madgpu context of the same size / spacing / device, configured by the caller as DiffusionStep does
 * (number_of_steps = DiffusionIterations, iterations_per_grid, cycle, tolerance, time step, max_cycles = 100).
 * in / out: HOST pointers of pixel type in_type / out_type (MADGPU_PIX_*).  solver_stats (may be NULL) receives the
 * statistics of the last DiffusionStep. */
int madved_run(madved_ctx *ctx, madgpu_ctx *solver, int32_t in_type, const void *in, int32_t out_type, void *out, const double *scales,
               int32_t nscales, int32_t iterations, madgpu_stats *solver_stats);

#ifdef __cplusplus
}
#endif
#endif /* MADVED_H */
