// ved_host_harness.cpp -- TEST INFRASTRUCTURE.  Compiles multigridanisotropicdiffusion_b200/csrc/ved_math.h -- the arithmetic
// the CUDA kernels of ved.cu execute -- for the HOST, so that the `-m "not gpu"` suite can check it against the oracle
// (oracle/ved_oracle.c) without a GPU: coefficient set-up, the line recursion with its fp32 intermediate storage, the
// pass structure of the separable Hessian (same order and buffers as hessian() in ved.cu), the eigen-solver, the
// vesselness function and the per-voxel update.  Built by tests/test_cpu_ved.py with g++ into tests/_build/.
// The product never loads this file; the CUDA path never runs on the CPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../multigridanisotropicdiffusion_b200/csrc/ved_math.h"

extern "C" {

void vh_rg_setup(double sigma, double spacing, int order, int normalize, double* out14)
{
  ved::RgCoefs c;
  ved::rg_setup(sigma, spacing, order, normalize != 0, c);
  std::memcpy(out14, &c, sizeof c);
}

// one line, one filter; fp32 in / out like the kernels
void vh_rg_line(double sigma, double spacing, int order, int normalize, const float* x, int n, float* y, double scale)
{
  ved::RgCoefs c;
  ved::rg_setup(sigma, spacing, order, normalize != 0, c);
  float* out[1] = {y};
  ved::rg_line<1>(x, 1, n, &c, out, &scale);
}

// hessian() of ved.cu on the host: n = (nx, ny, nz); image nvox floats; H six planes of nvox floats
void vh_hessian(const int* n, const double* h, double sigma, const float* image, float* H)
{
  const long long nx = n[0], ny = n[1], nz = n[2], nvox = nx * ny * nz;
  ved::RgCoefs c[3][3];
  for (int ax = 0; ax < 3; ++ax)
    for (int o = 0; o < 3; ++o) ved::rg_setup(sigma, h[ax], o, true, c[ax][o]);
  std::vector<std::vector<float> > W(12, std::vector<float>(nvox));
  const double one[3] = {1.0, 1.0, 1.0};
  {  // x pass: image -> G0x, G1x, G2x
    for (long long row = 0; row < ny * nz; ++row) {
      float* out[3] = {W[0].data() + row * nx, W[1].data() + row * nx, W[2].data() + row * nx};
      ved::rg_line<3>(image + row * nx, 1, (int)nx, c[0], out, one);
    }
  }
  auto lines = [&](int axis, const float* in, int K, const ved::RgCoefs* cc, float* const* outs, const double* scale) {
    const long long nlines = axis == 1 ? nx * nz : nx * ny, inner = axis == 1 ? nx : nx * ny, stride = axis == 1 ? nx : nx * ny;
    const int len = (int)(axis == 1 ? ny : nz);
    for (long long t = 0; t < nlines; ++t) {
      const long long base = (t / inner) * (nx * ny) + (t % inner);
      float* o[3];
      for (int k = 0; k < K; ++k) o[k] = outs[k] + base;
      if (K == 3) ved::rg_line<3>(in + base, stride, len, cc, o, scale);
      else if (K == 2) ved::rg_line<2>(in + base, stride, len, cc, o, scale);
      else ved::rg_line<1>(in + base, stride, len, cc, o, scale);
    }
  };
  float *Pxx = W[3].data(), *Pxy = W[4].data(), *Pxz = W[5].data(), *Pyy = W[6].data(), *Pyz = W[7].data(), *Pzz = W[8].data();
  {  // y pass
    float* o3[3] = {Pzz, Pyz, Pyy};
    lines(1, W[0].data(), 3, c[1], o3, one);
    float* o2[2] = {Pxz, Pxy};
    lines(1, W[1].data(), 2, c[1], o2, one);
    float* o1[1] = {Pxx};
    lines(1, W[2].data(), 1, c[1], o1, one);
  }
  struct Z { const float* in; float* out; int order; double factor; };
  const Z z[6] = {{Pxx, W[0].data(), 0, h[0] * h[0]}, {Pxy, W[1].data(), 0, h[0] * h[1]}, {Pxz, W[2].data(), 1, h[0] * h[2]},
                  {Pyy, W[9].data(), 0, h[1] * h[1]}, {Pyz, W[10].data(), 1, h[1] * h[2]}, {Pzz, W[11].data(), 2, h[2] * h[2]}};
  for (int k = 0; k < 6; ++k) {
    const double scale = 1.0 / z[k].factor;
    float* o1[1] = {z[k].out};
    lines(2, z[k].in, 1, &c[2][z[k].order], o1, &scale);
    std::memcpy(H + k * nvox, z[k].out, sizeof(float) * nvox);
  }
}

void vh_eig3_top(const double* h6, double* w, double* t) { ved::eig3_top(h6, w, t); }

double vh_vesselness(const double* e, double alpha, double beta, double gamma)
{
  const ved::Params P = {alpha, beta, gamma, 0.01, 5.0, 10.0};
  return ved::vesselness(e, P);
}

// k_ved_update on the host.  soa != 0: H is six fp32 planes; else hessian_aos holds 6 doubles per voxel.
// response nvox doubles, T six fp32 planes.
void vh_update(long long nvox, int soa, const float* H, const double* hessian_aos, int first, const double* params6, double* response, float* T)
{
  const ved::Params P = {params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
  for (long long v = 0; v < nvox; ++v) {
    double h[6], t[6];
    for (int k = 0; k < 6; ++k) h[k] = soa ? (double)H[k * nvox + v] : hessian_aos[v * 6 + k];
    double resp = first ? 0.0 : response[v];
    if (ved::update_voxel(h, first != 0, P, resp, t)) {
      response[v] = resp;
      for (int k = 0; k < 6; ++k) T[k * nvox + v] = (float)t[k];
    }
  }
}

}  // extern "C"
