// Standalone potential-energy kernel: Amber bonded terms + NonbondedForce(CutoffNonPeriodic,
// reaction field) + 1-4 exceptions + GBSA-OBC, one thread block per conformation, fp64 math.
//
// Replaces the OpenMM call behind OpenmmPotentialEnergyTorch.forward
// (utils/openmm/openmm_bridge.py:281-294 -> :170-249 -> bgflow -> OpenMM 7.7) for the systems
// of simulation/md.py:149-173.  OpenMM is a third-party dependency that is absent here
// (openmm==7.7, timewarp-environment.yml:22): the functional forms below restate OpenMM's
// documented/Reference-platform algorithms (HarmonicBondForce, HarmonicAngleForce,
// PeriodicTorsionForce, NonbondedForce with CutoffNonPeriodic, GBSAOBCForce / ReferenceObc).
// PARITY UNPINNED against the reference's golden energies (no force-field parameter files on
// this machine); pinned against oracle/energy_oracle.py (independent fp64 numpy restatement).
#include "common.cuh"
#include <curand_kernel.h>

namespace tw {

struct Vec3 {
  double x, y, z;
};
__device__ __forceinline__ Vec3 sub(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int ENERGY_THREADS = 128;

__device__ __forceinline__ void add3(double* g, int i, const Vec3& v, double w) {  // g[i] += w * v  (shared-memory fp64 atomics)
  atomicAdd(&g[3 * i], w * v.x);
  atomicAdd(&g[3 * i + 1], w * v.y);
  atomicAdd(&g[3 * i + 2], w * v.z);
}

// Energy terms (and, kForces, dE/dx in `grad`) of the conformation held in shared memory (`pos`); every thread of the block
// calls it; e[5] = this thread's partial sums of bond, angle, torsion, nonbonded(+exceptions), GB/SA.  `grad`, `dEdB`, `dBds`
// are contiguous (5N doubles) and are zeroed here.
template <bool kForces>
__device__ __forceinline__ void energy_terms(const tw_energy_system& s, const Vec3* pos, double* born, double* grad, double* dEdB,
                                             double* dBds, double (&e)[5]) {
  const int N = s.n_atoms;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (kForces) {
    for (int i = tid; i < 5 * N; i += nt) grad[i] = 0.0;  // grad, dEdB, dBds are contiguous
    __syncthreads();
  }
  double e_bond = 0, e_angle = 0, e_tors = 0, e_nb = 0, e_gb = 0;

  // HarmonicBondForce: 1/2 k (r - r0)^2
  for (int i = tid; i < s.n_bonds; i += nt) {
    Vec3 d = sub(pos[s.bond_idx[2 * i]], pos[s.bond_idx[2 * i + 1]]);
    double r = sqrt(dot(d, d)), dr = r - (double)s.bond_param[2 * i];
    e_bond += 0.5 * (double)s.bond_param[2 * i + 1] * dr * dr;
    if (kForces) {
      const double f = (double)s.bond_param[2 * i + 1] * dr / r;
      add3(grad, s.bond_idx[2 * i], d, f);
      add3(grad, s.bond_idx[2 * i + 1], d, -f);
    }
  }
  // HarmonicAngleForce: 1/2 k (theta - theta0)^2
  for (int i = tid; i < s.n_angles; i += nt) {
    Vec3 pj = pos[s.angle_idx[3 * i + 1]];
    Vec3 a = sub(pos[s.angle_idx[3 * i]], pj), c = sub(pos[s.angle_idx[3 * i + 2]], pj);
    const double la2 = dot(a, a), lc2 = dot(c, c), inv = 1.0 / sqrt(la2 * lc2);
    double cs = dot(a, c) * inv;
    cs = fmin(1.0, fmax(-1.0, cs));
    double dth = acos(cs) - (double)s.angle_param[2 * i];
    e_angle += 0.5 * (double)s.angle_param[2 * i + 1] * dth * dth;
    if (kForces) {  // dE/dcos = -k (theta - theta0) / sin(theta)
      const double de = -(double)s.angle_param[2 * i + 1] * dth / fmax(sqrt(1.0 - cs * cs), 1e-12);
      const Vec3 ga = {de * (c.x * inv - cs * a.x / la2), de * (c.y * inv - cs * a.y / la2), de * (c.z * inv - cs * a.z / la2)};
      const Vec3 gc = {de * (a.x * inv - cs * c.x / lc2), de * (a.y * inv - cs * c.y / lc2), de * (a.z * inv - cs * c.z / lc2)};
      add3(grad, s.angle_idx[3 * i], ga, 1.0);
      add3(grad, s.angle_idx[3 * i + 2], gc, 1.0);
      add3(grad, s.angle_idx[3 * i + 1], {ga.x + gc.x, ga.y + gc.y, ga.z + gc.z}, -1.0);
    }
  }
  // PeriodicTorsionForce: k (1 + cos(n phi - phase)); phi with the IUPAC sign (OpenMM Reference:
  // cross products of (p0-p1),(p2-p1),(p2-p3); sign from (p0-p1).cross2)
  for (int i = tid; i < s.n_torsions; i += nt) {
    Vec3 p0 = pos[s.torsion_idx[4 * i]], p1 = pos[s.torsion_idx[4 * i + 1]], p2 = pos[s.torsion_idx[4 * i + 2]],
         p3 = pos[s.torsion_idx[4 * i + 3]];
    Vec3 d0 = sub(p0, p1), d1 = sub(p2, p1), d2 = sub(p2, p3);
    Vec3 c1 = cross(d0, d1), c2 = cross(d1, d2);
    double cs = dot(c1, c2) / sqrt(dot(c1, c1) * dot(c2, c2));
    cs = fmin(1.0, fmax(-1.0, cs));
    double phi = acos(cs);
    if (dot(d0, c2) < 0) phi = -phi;
    e_tors += (double)s.torsion_param[3 * i + 2] *
              (1.0 + cos((double)s.torsion_param[3 * i] * phi - (double)s.torsion_param[3 * i + 1]));
    if (kForces) {  // dphi/dp (Blondel & Karplus) with F = d0, G = p1 - p2 = -d1, H = p3 - p2 = -d2, A = c1, B = c2
      const double n = (double)s.torsion_param[3 * i];
      const double dE = -(double)s.torsion_param[3 * i + 2] * n * sin(n * phi - (double)s.torsion_param[3 * i + 1]);
      const double lG = sqrt(dot(d1, d1)), A2 = dot(c1, c1), B2 = dot(c2, c2);
      const double fg = -dot(d0, d1), hg = dot(d2, d1);
      const double wa0 = lG / A2, wb3 = -lG / B2;
      const double wa1 = -lG / A2 - fg / (A2 * lG), wb1 = hg / (B2 * lG);
      const Vec3 g0 = {wa0 * c1.x, wa0 * c1.y, wa0 * c1.z};
      const Vec3 g3 = {wb3 * c2.x, wb3 * c2.y, wb3 * c2.z};
      const Vec3 g1 = {wa1 * c1.x + wb1 * c2.x, wa1 * c1.y + wb1 * c2.y, wa1 * c1.z + wb1 * c2.z};
      const Vec3 g2 = {-(g0.x + g1.x + g3.x), -(g0.y + g1.y + g3.y), -(g0.z + g1.z + g3.z)};
      add3(grad, s.torsion_idx[4 * i], g0, dE);
      add3(grad, s.torsion_idx[4 * i + 1], g1, dE);
      add3(grad, s.torsion_idx[4 * i + 2], g2, dE);
      add3(grad, s.torsion_idx[4 * i + 3], g3, dE);
    }
  }
  // 1-4 exceptions: plain Coulomb + LJ with the exception parameters, no cutoff
  for (int i = tid; i < s.n_exceptions; i += nt) {
    Vec3 d = sub(pos[s.exception_idx[2 * i]], pos[s.exception_idx[2 * i + 1]]);
    double r2 = dot(d, d), inv_r = 1.0 / sqrt(r2);
    double sig = (double)s.exception_param[3 * i + 1], eps = (double)s.exception_param[3 * i + 2];
    double sr2 = sig * sig / r2, sr6 = sr2 * sr2 * sr2;
    e_nb += 4.0 * eps * (sr6 * sr6 - sr6) + s.one_4pi_eps0 * (double)s.exception_param[3 * i] * inv_r;
    if (kForces) {  // (dE/dr) / r
      const double f = ((-48.0 * eps * sr6 * sr6 + 24.0 * eps * sr6) - s.one_4pi_eps0 * (double)s.exception_param[3 * i] * inv_r) / r2;
      add3(grad, s.exception_idx[2 * i], d, f);
      add3(grad, s.exception_idx[2 * i + 1], d, -f);
    }
  }
  // NonbondedForce, CutoffNonPeriodic: LJ (Lorentz-Berthelot) truncated at the cutoff + Coulomb
  // with reaction field  qq (1/r + k_rf r^2 - c_rf).
  const bool use_cut = s.cutoff > 0;
  const double rc = s.cutoff;
  const double krf = use_cut ? (1.0 / (rc * rc * rc)) * (s.reaction_field_eps - 1.0) / (2.0 * s.reaction_field_eps + 1.0) : 0.0;
  const double crf = use_cut ? (1.0 / rc) * (3.0 * s.reaction_field_eps) / (2.0 * s.reaction_field_eps + 1.0) : 0.0;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int i = warp; i < N; i += nwarps) {
    const Vec3 pi = pos[i];
    const double qi = (double)s.charge[i], si = (double)s.sigma[i], ei = (double)s.epsilon[i];
    for (int j = i + 1 + lane; j < N; j += 32) {
      if (s.excluded[(size_t)i * N + j]) continue;
      Vec3 d = sub(pi, pos[j]);
      double r2 = dot(d, d), r = sqrt(r2);
      if (use_cut && r > rc) continue;
      double sig = 0.5 * (si + (double)s.sigma[j]), eps = sqrt(ei * (double)s.epsilon[j]);
      double sr2 = sig * sig / r2, sr6 = sr2 * sr2 * sr2;
      e_nb += 4.0 * eps * (sr6 * sr6 - sr6) + s.one_4pi_eps0 * qi * (double)s.charge[j] * (1.0 / r + krf * r2 - crf);
      if (kForces) {
        const double f = (-48.0 * eps * sr6 * sr6 + 24.0 * eps * sr6) / r2 + s.one_4pi_eps0 * qi * (double)s.charge[j] * (-1.0 / (r2 * r) + 2.0 * krf);
        add3(grad, i, d, f);
        add3(grad, j, d, -f);
      }
    }
  }

  if (s.use_gb) {
    // ---- Born radii (ReferenceObc::computeBornRadii) ----
    for (int i = warp; i < N; i += nwarps) {
      const Vec3 pi = pos[i];
      const double ri = (double)s.gb_radius[i], ori = ri - s.gb_offset;
      double sum = 0;
      for (int j = lane; j < N; j += 32) {
        if (j == i) continue;
        Vec3 d = sub(pi, pos[j]);
        double r = sqrt(dot(d, d));
        if (use_cut && r > rc) continue;
        double srj = ((double)s.gb_radius[j] - s.gb_offset) * (double)s.gb_scale[j];
        double rsr = r + srj;
        if (ori < rsr) {
          double rinv = 1.0 / r;
          double adiff = fabs(r - srj);
          double l_ij = 1.0 / (ori > adiff ? ori : adiff);
          double u_ij = 1.0 / rsr;
          double l2 = l_ij * l_ij, u2 = u_ij * u_ij;
          double ratio = log(u_ij / l_ij);
          double term = l_ij - u_ij + 0.25 * r * (u2 - l2) + 0.5 * rinv * ratio + 0.25 * srj * srj * rinv * (l2 - u2);
          if (ori < (srj - r)) term += 2.0 * (1.0 / ori - l_ij);
          sum += term;
        }
      }
      sum = warp_sum_d(sum);
      if (lane == 0) {
        sum *= 0.5 * ori;
        double s2 = sum * sum, s3 = sum * s2;
        double th = tanh(s.gb_alpha * sum - s.gb_beta * s2 + s.gb_gamma * s3);
        born[i] = 1.0 / (1.0 / ori - th / ri);
        if (kForces)
          dBds[i] = born[i] * born[i] * (1.0 - th * th) * (s.gb_alpha - 2.0 * s.gb_beta * sum + 3.0 * s.gb_gamma * s2) / ri * 0.5 * ori;
      }
    }
    __syncthreads();
    // ---- ACE surface-area term ----
    if (s.surface_area_energy != 0.0) {
      for (int i = tid; i < N; i += nt) {
        if (born[i] > 0) {
          double ri = (double)s.gb_radius[i], rr = ri + 0.14, q = ri / born[i];
          double q2 = q * q;
          const double e_sa = s.surface_area_energy * rr * rr * q2 * q2 * q2;
          e_gb += e_sa;
          if (kForces) atomicAdd(&dEdB[i], -6.0 * e_sa / born[i]);
        }
      }
    }
    // ---- polarisation energy (ReferenceObc::computeBornEnergyForces), pairs j >= i ----
    const double pre = (s.solute_dielectric != 0 && s.solvent_dielectric != 0)
                           ? -s.one_4pi_eps0 * (1.0 / s.solute_dielectric - 1.0 / s.solvent_dielectric)
                           : 0.0;
    for (int i = warp; i < N; i += nwarps) {
      const Vec3 pi = pos[i];
      const double qi = pre * (double)s.charge[i], bi = born[i];
      for (int j = i + lane; j < N; j += 32) {
        Vec3 d = sub(pi, pos[j]);
        double r2 = dot(d, d);
        if (use_cut && r2 > rc * rc) continue;
        double a2 = bi * born[j];
        const double ex = exp(-r2 / (4.0 * a2));
        const double D = r2 + a2 * ex;
        double den = sqrt(D);
        double g = qi * (double)s.charge[j] / den;
        if (kForces) {  // E = c / sqrt(D):  dD/dr^2 = 1 - ex/4,  dD/d(B_i B_j) = ex (1 + r^2 / (4 B_i B_j))
          const double cc = qi * (double)s.charge[j] * (j == i ? 0.5 : 1.0);
          const double dEdD = -0.5 * cc / (D * den);
          const double dDa2 = ex * (1.0 + r2 / (4.0 * a2));
          if (j != i) {
            const double f = dEdD * (1.0 - 0.25 * ex) * 2.0;
            add3(grad, i, d, f);
            add3(grad, j, d, -f);
            atomicAdd(&dEdB[i], dEdD * dDa2 * born[j]);
            atomicAdd(&dEdB[j], dEdD * dDa2 * bi);
          } else {
            atomicAdd(&dEdB[i], dEdD * dDa2 * 2.0 * bi);
          }
        }
        if (j != i) {
          if (use_cut) g -= qi * (double)s.charge[j] / rc;
        } else {
          g *= 0.5;
        }
        e_gb += g;
      }
    }
    if (kForces) {
      // ---- chain rule through the Born radii: s_i = 0.5 (r_i - offset) sum_j term(r_ij) ----
      __syncthreads();
      for (int i = warp; i < N; i += nwarps) {
        const Vec3 pi = pos[i];
        const double ori = (double)s.gb_radius[i] - s.gb_offset;
        const double w = dEdB[i] * dBds[i];
        for (int j = lane; j < N; j += 32) {
          if (j == i) continue;
          Vec3 d = sub(pi, pos[j]);
          double r = sqrt(dot(d, d));
          if (use_cut && r > rc) continue;
          double srj = ((double)s.gb_radius[j] - s.gb_offset) * (double)s.gb_scale[j];
          double rsr = r + srj;
          if (ori < rsr) {
            const double ad = fabs(r - srj);
            double l, dl;
            if (ori > ad) l = 1.0 / ori, dl = 0.0;
            else l = 1.0 / ad, dl = -l * l * (r >= srj ? 1.0 : -1.0);
            const double u = 1.0 / rsr, du = -u * u;
            const double rinv = 1.0 / r, l2 = l * l, u2 = u * u;
            double dterm = dl - du + 0.25 * (u2 - l2) + 0.5 * r * (u * du - l * dl) - 0.5 * rinv * rinv * log(u / l) +
                           0.5 * rinv * (du / u - dl / l) - 0.25 * srj * srj * rinv * rinv * (l2 - u2) +
                           0.5 * srj * srj * rinv * (l * dl - u * du);
            if (ori < (srj - r)) dterm -= 2.0 * dl;
            const double f = w * dterm * rinv;
            add3(grad, i, d, f);
            add3(grad, j, d, -f);
          }
        }
      }
    }
  }

  e[0] = e_bond, e[1] = e_angle, e[2] = e_tors, e[3] = e_nb, e[4] = e_gb;
  if (kForces) __syncthreads();  // every contribution to grad has landed
}

// kForces: also accumulate the analytic gradient dE/dx of every term in shared memory (fp64) and write the forces
// F = -dE/dx (what OpenMM returns and bgflow feeds back as the gradient, openmm_bridge.py:56-60).  The GB term needs the
// chain rule through the Born radii: dE/dB_i is accumulated with the pair energies, then a second pair loop applies
// dB_i/dr_ij (derivative of the OBC descreening integral).  Formulas checked against central differences of the fp64
// oracle (tests/test_gpu_energy_mh.py::test_forces_match_finite_differences).
template <bool kForces>
__global__ void __launch_bounds__(ENERGY_THREADS) k_energy(tw_energy_system s, const float* __restrict__ coords,
                                                           float* __restrict__ out_energy, float* __restrict__ out_terms,
                                                           float* __restrict__ out_forces) {
  extern __shared__ double sm[];
  const int N = s.n_atoms;
  Vec3* pos = reinterpret_cast<Vec3*>(sm);            // [N]
  double* born = sm + 3 * (size_t)N;                  // [N]
  double* grad = born + N;                            // [3N]  dE/dx          (kForces)
  double* dEdB = grad + 3 * (size_t)N;                // [N]   dE/dBorn_i     (kForces)
  double* dBds = dEdB + N;                            // [N]   dBorn_i/ds_i * 0.5 * (r_i - offset)   (kForces)
  __shared__ double red[ENERGY_THREADS / 32][5];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int64_t b = blockIdx.x;
  const float* cb = coords + b * (int64_t)N * 3;
  for (int i = tid; i < N; i += nt) pos[i] = {(double)cb[i * 3], (double)cb[i * 3 + 1], (double)cb[i * 3 + 2]};
  __syncthreads();
  double v[5];
  energy_terms<kForces>(s, pos, born, grad, dEdB, dBds, v);
#pragma unroll
  for (int k = 0; k < 5; k++) {
    double t = warp_sum_d(v[k]);
    if (lane == 0) red[warp][k] = t;
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0, terms[5];
    for (int k = 0; k < 5; k++) {
      terms[k] = 0;
      for (int w = 0; w < nwarps; w++) terms[k] += red[w][k];
      tot += terms[k];
      if (out_terms) out_terms[b * 5 + k] = (float)terms[k];
    }
    out_energy[b] = (float)tot;
  }
  if (kForces) {
    float* fb = out_forces + b * (int64_t)N * 3;
    for (int i = tid; i < 3 * N; i += nt) fb[i] = (float)(-grad[i]);
  }
}


// ------------------------------------------------------------------------------------------
// OpenMM integrator steps on the device: the `sim.step(num_steps)` of openmm_step (utils/evaluation_utils.py:439-464) for the
// integrators of simulation/md.py:116-123, constraints=None (md.py:171,180).  One thread block per conformation; positions,
// velocities and forces stay in shared memory (fp64) for all n_steps; the force evaluation is energy_terms<true>.
//   kMiddle = false, LangevinIntegrator (OpenMM ReferenceStochasticDynamics / langevin.cu):
//       v' = a v + (1 - a)/gamma * F(x)/m + sqrt(kT (1 - a^2) / m) * xi,   x' = x + dt v',      a = exp(-gamma dt)
//     (gamma = 0: (1 - a)/gamma -> dt).  PINNED by the reference's trajectory fixtures: consecutive frames of
//     simulation/testdata/implicit-2olx-traj*-arrays.npz satisfy x' = x + dt v' to fp32 round-off and the implied xi has
//     unit variance (tests/test_md_oracle.py).
//   kMiddle = true, LangevinMiddleIntegrator (OpenMM >= 7.5, LangevinMiddleIntegrator docs / langevinMiddle.cu):
//       v += dt F(x)/m;  x += dt/2 v;  v = a v + sqrt(kT (1 - a^2) / m) * xi;  x += dt/2 v
//     restated from OpenMM's documentation (no fixture of the reference was generated with it that contains consecutive frames).
// xi: the caller's standard normals noise[step, b, atom, 3] (torch RNG), or, noise == NULL, Philox4x32-10 keyed by
// (seed, conformation * blockDim + thread) at `offset`.
template <bool kMiddle>
__global__ void __launch_bounds__(ENERGY_THREADS) k_langevin(tw_energy_system s, float* __restrict__ coords, float* __restrict__ velocs,
                                                             const float* __restrict__ masses, const float* __restrict__ noise,
                                                             int64_t B, int n_steps, double dt, double vscale, double fscale,
                                                             double kT, unsigned long long seed, unsigned long long offset) {
  extern __shared__ double sm[];
  const int N = s.n_atoms;
  Vec3* pos = reinterpret_cast<Vec3*>(sm);  // [N]
  double* born = sm + 3 * (size_t)N;        // [N]
  double* grad = born + N;                  // [3N]
  double* dEdB = grad + 3 * (size_t)N;      // [N]
  double* dBds = dEdB + N;                  // [N]
  double* vel = dBds + N;                   // [3N]
  double* invm = vel + 3 * (size_t)N;       // [N]
  double* x = sm;                           // pos as a flat [3N] array
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t b = blockIdx.x;
  float* cb = coords + b * (int64_t)N * 3;
  float* vb = velocs + b * (int64_t)N * 3;
  for (int i = tid; i < 3 * N; i += nt) x[i] = (double)cb[i], vel[i] = (double)vb[i];
  for (int i = tid; i < N; i += nt) invm[i] = 1.0 / (double)masses[i];
  curandStatePhilox4_32_10_t rng;
  if (!noise) curand_init(seed, (unsigned long long)b * nt + tid, offset, &rng);
  const double nscale = sqrt(kT * (1.0 - vscale * vscale));
  __syncthreads();
  double e[5];
  for (int step = 0; step < n_steps; step++) {
    energy_terms<true>(s, pos, born, grad, dEdB, dBds, e);  // grad = dE/dx = -F; ends with a block barrier
    const float* nz = noise ? noise + ((int64_t)step * B + b) * (int64_t)N * 3 : nullptr;
    for (int i = tid; i < 3 * N; i += nt) {
      const double im = invm[i / 3];
      const double xi = nz ? (double)nz[i] : (double)curand_normal(&rng);
      double v = vel[i];
      if (kMiddle) {
        v -= dt * grad[i] * im;
        double xx = x[i] + 0.5 * dt * v;
        v = vscale * v + nscale * sqrt(im) * xi;
        x[i] = xx + 0.5 * dt * v;
      } else {
        v = vscale * v - fscale * grad[i] * im + nscale * sqrt(im) * xi;
        x[i] += dt * v;
      }
      vel[i] = v;
    }
    __syncthreads();
  }
  for (int i = tid; i < 3 * N; i += nt) cb[i] = (float)x[i], vb[i] = (float)vel[i];
}

}  // namespace tw

using namespace tw;

extern "C" int tw_peptide_energy(const tw_energy_system* sys, const float* coords, int64_t B, float* out_energy,
                                 float* out_forces, float* out_terms, void* stream) {
  TW_CHECK_ARG(sys != nullptr, "NULL system");
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(coords && out_energy, "NULL pointer");
  TW_CHECK_ARG(sys->n_atoms >= 1 && sys->n_atoms <= 4096, "n_atoms out of range (1..4096)");
  TW_CHECK_ARG(B >= 0 && B <= 2147483647LL, "bad batch size");
  TW_CHECK_ARG(sys->n_bonds == 0 || (sys->bond_idx && sys->bond_param), "bond arrays missing");
  TW_CHECK_ARG(sys->n_angles == 0 || (sys->angle_idx && sys->angle_param), "angle arrays missing");
  TW_CHECK_ARG(sys->n_torsions == 0 || (sys->torsion_idx && sys->torsion_param), "torsion arrays missing");
  TW_CHECK_ARG(sys->n_exceptions == 0 || (sys->exception_idx && sys->exception_param), "exception arrays missing");
  TW_CHECK_ARG(sys->charge && sys->sigma && sys->epsilon && sys->excluded, "nonbonded arrays missing");
  TW_CHECK_ARG(!sys->use_gb || (sys->gb_radius && sys->gb_scale), "GB arrays missing");
  TW_CHECK_ARG(!out_forces || sys->n_atoms <= 2048, "forces: n_atoms out of range (1..2048)");
  size_t smem = (size_t)sys->n_atoms * (out_forces ? 9 : 4) * sizeof(double);
  static DeviceOnce attr_set;
  if (smem > 48 * 1024 && !attr_set.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_energy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 4 * (int)sizeof(double)));
    TW_CUDA(cudaFuncSetAttribute(k_energy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 9 * (int)sizeof(double)));
    attr_set.mark();
  }
  {
    ProfScope prof(PROF_ENERGY, (cudaStream_t)stream);
    if (out_forces)
      k_energy<true><<<(unsigned)B, ENERGY_THREADS, smem, (cudaStream_t)stream>>>(*sys, coords, out_energy, out_terms, out_forces);
    else
      k_energy<false><<<(unsigned)B, ENERGY_THREADS, smem, (cudaStream_t)stream>>>(*sys, coords, out_energy, out_terms, nullptr);
  }
  TW_LAUNCH_CHECK();
  return TW_OK;
}

static int check_system(const tw_energy_system* sys) {
  TW_CHECK_ARG(sys != nullptr, "NULL system");
  TW_CHECK_ARG(sys->n_atoms >= 1 && sys->n_atoms <= 4096, "n_atoms out of range (1..4096)");
  TW_CHECK_ARG(sys->n_bonds == 0 || (sys->bond_idx && sys->bond_param), "bond arrays missing");
  TW_CHECK_ARG(sys->n_angles == 0 || (sys->angle_idx && sys->angle_param), "angle arrays missing");
  TW_CHECK_ARG(sys->n_torsions == 0 || (sys->torsion_idx && sys->torsion_param), "torsion arrays missing");
  TW_CHECK_ARG(sys->n_exceptions == 0 || (sys->exception_idx && sys->exception_param), "exception arrays missing");
  TW_CHECK_ARG(sys->charge && sys->sigma && sys->epsilon && sys->excluded, "nonbonded arrays missing");
  TW_CHECK_ARG(!sys->use_gb || (sys->gb_radius && sys->gb_scale), "GB arrays missing");
  return TW_OK;
}

extern "C" int tw_langevin_steps(const tw_energy_system* sys, float* coords, float* velocs, const float* masses, int64_t B,
                                 int32_t n_steps, int32_t integrator, double timestep, double friction, double kT,
                                 const float* noise, uint64_t seed, uint64_t offset, void* stream) {
  TW_TRY(check_system(sys));
  TW_CHECK_ARG(integrator == TW_INTEGRATOR_LANGEVIN || integrator == TW_INTEGRATOR_LANGEVIN_MIDDLE, "unknown integrator");
  TW_CHECK_ARG(n_steps >= 0 && timestep > 0 && friction >= 0 && kT >= 0, "bad integrator parameters");
  if (B == 0 || n_steps == 0) return TW_OK;
  TW_CHECK_ARG(coords && velocs && masses, "NULL pointer");
  TW_CHECK_ARG(B >= 0 && B <= 2147483647LL, "bad batch size");
  TW_CHECK_ARG(sys->n_atoms <= 1536, "integrator: n_atoms out of range (1..1536)");
  const size_t smem = (size_t)sys->n_atoms * 13 * sizeof(double);
  static DeviceOnce attr_set;
  if (smem > 48 * 1024 && !attr_set.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_langevin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 13 * (int)sizeof(double)));
    TW_CUDA(cudaFuncSetAttribute(k_langevin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 13 * (int)sizeof(double)));
    attr_set.mark();
  }
  const double vscale = exp(-timestep * friction);
  const double fscale = friction == 0 ? timestep : (1.0 - vscale) / friction;
  if (integrator == TW_INTEGRATOR_LANGEVIN_MIDDLE)
    k_langevin<true><<<(unsigned)B, ENERGY_THREADS, smem, (cudaStream_t)stream>>>(*sys, coords, velocs, masses, noise, B, n_steps, timestep,
                                                                                 vscale, fscale, kT, seed, offset);
  else
    k_langevin<false><<<(unsigned)B, ENERGY_THREADS, smem, (cudaStream_t)stream>>>(*sys, coords, velocs, masses, noise, B, n_steps, timestep,
                                                                                  vscale, fscale, kT, seed, offset);
  TW_LAUNCH_CHECK();
  return TW_OK;
}
