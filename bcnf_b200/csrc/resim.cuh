// Batched re-simulation of sampled parameter sets: the reference integrates one trajectory per
// scipy.integrate.odeint call in a 32-process pool (src/bcnf/simulation/resimulation.py:21-59, physics.py:53-165;
// 3 min 47 s per 10^6 trajectories, notebooks/resimulation.ipynb:253).  Here one thread integrates one trajectory.
//
//   dv/dt = g - g rho (4/3) pi r^3 / m - (b / 2m) (v_i^3 / |v| - w_i^3 / |w|) + a            physics.py:48 (element-wise cubes)
//   x_0 = x0 ;  x_i = x_(i-1) + v(t_i) dt                                                     physics.py:150-153
//   break_on_impact: first i with x_i,z < 0 -> x_i = x_(i-1) + v_i * (-x_(i-1),z / v_i,z), held for the rest   physics.py:156-162
//
// v(t_i) comes from classical RK4 in fp64 with `substeps` steps per output interval (the reference's LSODA runs at
// rtol = atol = 1.49e-8; 16 substeps of dt = 0.1 leave < 1e-9 of scale on these smooth right-hand sides).
#pragma once
#include <cuda_runtime.h>

namespace bcnf {

// parameter order of one row of `params`: the keyword order of physics_ODE_simulation (physics.py:53-72)
enum ResimParam { RP_X0 = 0, RP_V0 = 3, RP_G = 6, RP_W = 9, RP_B = 12, RP_M = 13, RP_RHO = 14, RP_R = 15, RP_A = 16, RP_COUNT = 19 };

struct ResimRhs {
  double c[3];      // g - g rho (4/3) pi r^3 / m + a + (b / 2m) w_i^3 / |w|   (constant part)
  double k;         // b / 2m
  __device__ __forceinline__ void operator()(const double (&v)[3], double (&dv)[3]) const {
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) dv[i] = c[i] - k * (v[i] * v[i] * v[i] / n);     // (0 / 0 = NaN at v = 0, as the reference)
  }
};

__global__ void resim_kernel(const double* __restrict__ params, long long n, int n_steps, double dt, int substeps,
                             int break_on_impact, double* __restrict__ x_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* p = params + t * RP_COUNT;
  const double kPi = 3.14159265358979323846;
  ResimRhs f;
  const double buoy = p[RP_RHO] * (4.0 / 3.0) * (kPi * p[RP_R] * p[RP_R] * p[RP_R]) / p[RP_M];
  f.k = 0.5 * p[RP_B] / p[RP_M];
  const double wn = sqrt(p[RP_W] * p[RP_W] + p[RP_W + 1] * p[RP_W + 1] + p[RP_W + 2] * p[RP_W + 2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double w = p[RP_W + i];
    f.c[i] = p[RP_G + i] - p[RP_G + i] * buoy + f.k * (w * w * w / wn) + p[RP_A + i];
  }
  double v[3] = {p[RP_V0], p[RP_V0 + 1], p[RP_V0 + 2]};
  double x[3] = {p[RP_X0], p[RP_X0 + 1], p[RP_X0 + 2]};
  double* out = x_out + t * (long long)n_steps * 3;
  out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
  const double h = dt / substeps;
  bool landed = false;
  for (int s = 1; s < n_steps; ++s) {
    if (!landed) {
      for (int u = 0; u < substeps; ++u) {
        double k1[3], k2[3], k3[3], k4[3], y[3];
        f(v, k1);
#pragma unroll
        for (int i = 0; i < 3; ++i) y[i] = v[i] + 0.5 * h * k1[i];
        f(y, k2);
#pragma unroll
        for (int i = 0; i < 3; ++i) y[i] = v[i] + 0.5 * h * k2[i];
        f(y, k3);
#pragma unroll
        for (int i = 0; i < 3; ++i) y[i] = v[i] + h * k3[i];
        f(y, k4);
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] += (h / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
      }
      double xn[3] = {x[0] + v[0] * dt, x[1] + v[1] * dt, x[2] + v[2] * dt};
      if (break_on_impact && xn[2] < 0.0) {
        const double ti = -x[2] / v[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) xn[i] = x[i] + v[i] * ti;
        landed = true;
      }
      x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
    }
    out[s * 3 + 0] = x[0]; out[s * 3 + 1] = x[1]; out[s * 3 + 2] = x[2];
  }
}

}  // namespace bcnf
