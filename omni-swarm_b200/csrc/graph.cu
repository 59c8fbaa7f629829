// graph.cu -- osb_solver: GPU pose-graph solve replacing the body of SwarmLocalizationSolver::solve_once
// (swarm_localization/src/swarm_localization_solver.cpp:1668-1725).
//
// Factors (swarm_localization/include/swarm_localization/swarm_localization_factors.hpp):
//   DistanceMeasurementFactor :203-224, RelativePoseFactor4d :226-271 (DeltaPose :139-149, pose_error_4d :52-61),
//   DroneDetection4dFactor :273-367 (PoseMulti :165-172, DeltaPose_Naive :153-160, unit_position_error* :73-103).
// The reference evaluates them with Ceres AutoDiff jets; here residuals and ANALYTIC Jacobians are evaluated by
// one thread per factor.  HuberLoss(1.0) is applied as Ceres' corrector does when
// rho'' <= 0: residual and Jacobian scaled by sqrt(rho'(s)).
//
// The solve is ONE persistent cooperative kernel: Levenberg-Marquardt outer loop (Ceres' LM step control: diagonal
// D = clip(diag(J^T J)), radius update by 1 - (2 rho - 1)^3, decrease factor doubling) and a block-Jacobi
// preconditioned conjugate gradient on the 4x4-block normal equations.  J^T J is never assembled off-diagonal:
// a factor thread computes t = Ja p_a + Jb p_b and the two 4-vectors Ja^T t, Jb^T t, and a node thread gathers the
// contributions from its own contiguous run of slots in a fixed order (no atomics: results are bit-reproducible
// run to run).  The problem (8 000 scalars, 12 000 factors for BASELINE config C5) is latency bound, not HBM bound:
// the Jacobians of a CTA's factor block stay in shared memory (SoA, conflict-free) for the whole solve, and when the
// factor list fits 16 CTAs the grid is launched as ONE thread-block cluster so that the 3 barriers per CG iteration
// are hardware cluster barriers instead of software grid barriers.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstring>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace osb {

constexpr int GS_THREADS = 256;           // 255 registers per thread: the CG fast path keeps ~50 doubles live per thread
constexpr int GS_KF = 3;                  // factors per thread on the fast path (16 CTAs x 256 threads x 3 >= 12 288 factors)
constexpr int GS_MAX_CLUSTER = 16;
constexpr int GS_SMEM_J_MAX = 200 * 1024;   // bytes of shared memory a CTA may spend on its Jacobian block
constexpr int GS_SMEM_DYN_MAX = 227 * 1024 - 2048;  // opt-in limit minus the kernel's static shared memory
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

__device__ __forceinline__ double normalize_angle(double a) {   // factors.hpp:34-40
  return a - kTwoPi * floor((a + kPi) / kTwoPi);
}

// residual (nr rows) and 4x4 Jacobian blocks (row-major, rows >= nr zero) of one factor, un-robustified.
// JAC = false: residual only (trial-point cost; identical arithmetic for r).
template <bool JAC>
__device__ __forceinline__ int linearize_factor_t(int type, const double* __restrict__ pa, const double* __restrict__ pb,
                                                  const double* __restrict__ pl, double r[4], double Ja[16], double Jb[16]) {
  if (JAC) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { Ja[i] = 0.0; Jb[i] = 0.0; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = 0.0;
  if (type == OSB_FACTOR_DISTANCE) {
    const double dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
    const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
    const double si = pl[1];
    r[0] = (nrm - pl[0]) * si;
    const double inv = si / nrm;
    Ja[0] = dx * inv; Ja[1] = dy * inv; Ja[2] = dz * inv;
    Jb[0] = -dx * inv; Jb[1] = -dy * inv; Jb[2] = -dz * inv;
    return 1;
  }
  if (type == OSB_FACTOR_RELPOSE) {
    double s, c;
    sincos(pa[3], &s, &c);
    const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
    double e[4];
    e[0] = pl[0] - (c * dx + s * dy);
    e[1] = pl[1] - (-s * dx + c * dy);
    e[2] = pl[2] - dz;
    e[3] = normalize_angle(pl[3] - normalize_angle(pb[3] - pa[3]));
    const double* S = pl + 4;
    // d est / d pose_a (columns x,y,z,yaw) and pose_b
    const double Ea[4][4] = {{-c, -s, 0.0, -s * dx + c * dy}, {s, -c, 0.0, -c * dx - s * dy}, {0.0, 0.0, -1.0, 0.0},
                             {0.0, 0.0, 0.0, -1.0}};
    const double Eb[4][4] = {{c, s, 0.0, 0.0}, {-s, c, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 0.0, 0.0, 1.0}};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double ri = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) ri += S[i * 4 + k] * e[k];
      r[i] = ri;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { a += S[i * 4 + k] * Ea[k][j]; b += S[i * 4 + k] * Eb[k][j]; }
        Ja[i * 4 + j] = -a; Jb[i * 4 + j] = -b;
      }
    }
    return 4;
  }
  // OSB_FACTOR_DETECTION
  const int flags = (int)pl[10];
  const double inv_dep = pl[9], ext_z = pl[11], sphere_std = pl[20], invdep_std = pl[21];
  double A[4], Bp[4];
  double Ga03 = 0.0, Ga13 = 0.0, Gb03 = 0.0, Gb13 = 0.0;   // d T' / d yaw of the raw pose (PoseMulti)
  if (flags & 2) {
    double s, c;
    sincos(pa[3], &s, &c);
    A[0] = pa[0] + c * pl[12] - s * pl[13]; A[1] = pa[1] + s * pl[12] + c * pl[13]; A[2] = pa[2] + pl[14];
    A[3] = normalize_angle(pa[3] + pl[15]);
    Ga03 = -s * pl[12] - c * pl[13]; Ga13 = c * pl[12] - s * pl[13];
    sincos(pb[3], &s, &c);
    Bp[0] = pb[0] + c * pl[16] - s * pl[17]; Bp[1] = pb[1] + s * pl[16] + c * pl[17]; Bp[2] = pb[2] + pl[18];
    Bp[3] = normalize_angle(pb[3] + pl[19]);
    Gb03 = -s * pl[16] - c * pl[17]; Gb13 = c * pl[16] - s * pl[17];
  } else {
    A[0] = pa[0]; A[1] = pa[1]; A[2] = pa[2] + ext_z; A[3] = pa[3];
    Bp[0] = pb[0]; Bp[1] = pb[1]; Bp[2] = pb[2]; Bp[3] = pb[3];
  }
  double s, c;
  sincos(A[3], &s, &c);
  const double dx = Bp[0] - A[0], dy = Bp[1] - A[1], dz = Bp[2] - A[2];
  const double rel[3] = {c * dx + s * dy, -s * dx + c * dy, dz};
  const double rho = 1.0 / sqrt(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
  // d rel / d A (4 cols), d rel / d B' (4 cols)
  const double RA[3][4] = {{-c, -s, 0.0, -s * dx + c * dy}, {s, -c, 0.0, -c * dx - s * dy}, {0.0, 0.0, -1.0, 0.0}};
  const double RB[3][4] = {{c, s, 0.0, 0.0}, {-s, c, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}};
  // chain through PoseMulti: columns 0..2 identity, column 3 gets + d rel/dT' * dT'/dyaw (yaw' = yaw + const)
  double DA[3][4], DB[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    DA[i][0] = RA[i][0]; DA[i][1] = RA[i][1]; DA[i][2] = RA[i][2];
    DA[i][3] = RA[i][3] + RA[i][0] * Ga03 + RA[i][1] * Ga13;
    DB[i][0] = RB[i][0]; DB[i][1] = RB[i][1]; DB[i][2] = RB[i][2];
    DB[i][3] = RB[i][3] + RB[i][0] * Gb03 + RB[i][1] * Gb13;
  }
  const int nr = (flags & 1) ? 3 : 2;
  double u[3], Jrel[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) u[i] = rel[i] * rho - pl[i];
  // dU = rho (I - rel rel^T rho^2)
  double dU[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) dU[i][j] = rho * ((i == j ? 1.0 : 0.0) - rel[i] * rel[j] * rho * rho);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double* Bt = pl + 3 + 3 * i;
    r[i] = (Bt[0] * u[0] + Bt[1] * u[1] + Bt[2] * u[2]) / sphere_std;
#pragma unroll
    for (int j = 0; j < 3; ++j) Jrel[i][j] = (Bt[0] * dU[0][j] + Bt[1] * dU[1][j] + Bt[2] * dU[2][j]) / sphere_std;
  }
  if (nr == 3) {
    r[2] = (inv_dep - rho) / invdep_std;
    const double r3 = rho * rho * rho / invdep_std;
    Jrel[2][0] = r3 * rel[0]; Jrel[2][1] = r3 * rel[1]; Jrel[2][2] = r3 * rel[2];
  } else {
    Jrel[2][0] = Jrel[2][1] = Jrel[2][2] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i >= nr) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Ja[i * 4 + j] = Jrel[i][0] * DA[0][j] + Jrel[i][1] * DA[1][j] + Jrel[i][2] * DA[2][j];
      Jb[i * 4 + j] = Jrel[i][0] * DB[0][j] + Jrel[i][1] * DB[1][j] + Jrel[i][2] * DB[2][j];
    }
  }
  return nr;
}

__device__ __forceinline__ int linearize_factor(int type, const double* __restrict__ pa, const double* __restrict__ pb,
                                                const double* __restrict__ pl, double r[4], double Ja[16], double Jb[16]) {
  return linearize_factor_t<true>(type, pa, pb, pl, r, Ja, Jb);
}
// |r|^2 only (un-robustified): the Jacobian arithmetic of the inlined body is dead code here
__device__ __forceinline__ double factor_sqnorm(int type, const double* __restrict__ pa, const double* __restrict__ pb,
                                                const double* __restrict__ pl) {
  double r[4], Ja[16], Jb[16];
  linearize_factor_t<false>(type, pa, pb, pl, r, Ja, Jb);
  return r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
}

// The inner (PCG) arithmetic type.  Levenberg-Marquardt only needs an INEXACT step (relative residual 1e-2), so
// everything inside the PCG -- Jacobian blocks, direction / residual / preconditioner vectors, contribution slots, the
// chain factorisation -- runs in fp32 when the requested pcg_tolerance allows (>= 1e-4); residuals, costs, the gradient
// J^T r, the diagonal, the poses and every LM decision stay fp64.  fp32 halves what bounds a CG iteration here: the
// 32-bit shuffles of the preconditioner sweeps, the bytes of every cross-thread gather through L2, the shared-memory
// footprint of the Jacobians (96 KB instead of 192 KB) and the live registers.  (It is NOT an fp64-throughput issue:
// scripts/microbench/fp64_rate.cu measures 62 DFMA lanes/clk/SM against 117 FFMA lanes/clk/SM on this B200.)  A numpy
// emulation of the same loop gives the same iteration counts and poses within 1e-8 of the all-fp64 solve;
// tests/test_gpu_solver.py checks it on the GPU.  T = double is kept for tight tolerances (parity tests).
struct SolverDev {
  int n, m;
  int fpc;                 // factors per CTA (contiguous block of the factor list)
  int use_cluster;         // 1: the whole grid is ONE thread-block cluster (hardware barrier); 0: cooperative grid sync
  int j_in_smem;           // 1: the CTA's Jacobians live in shared memory (SoA), 0: in `Jg`
  // graph
  const uint8_t* fixed; const int32_t* ftype; const int32_t* ia; const int32_t* ib; const uint8_t* huber;
  const double* payload;
  const int32_t* node_ptr;                 // CSR: node n owns contribution slots [node_ptr[n], node_ptr[n+1])
  const int32_t* slot_a; const int32_t* slot_b;   // slot of factor f's contribution to its node a / node b
  // state
  double* x[2];            // pose buffers (current / trial), [n][4]
  void* Jg;                // global Jacobian store, SoA [32][m] of T (used when the CTA block does not fit shared memory)
  double *g, *D, *Hnn;     // gradient, LM diagonal, diagonal Hessian blocks (fp64)
  void *Minv, *p, *z, *res, *Ap, *delta;   // PCG node vectors, element type T
  double* gs;              // gradient contribution slots [2m][4] (fp64, written at a linearisation)
  void* cs;                // PCG contribution slots [2m][4] of T
  double* hs;              // Hessian-diagonal-block slots [2m][16] (only touched at a re-linearisation)
  double* partial;         // [2][4][grid]
  osb_solve_options opt;
  osb_solve_summary* summary;
  double* poses_out;
  long long* dbg;          // [8] cycle counters of block 0 / thread 0 (profiling aid, see osb_solver_phase_cycles)
  // chain preconditioner (fast path only; see chain_apply)
  int use_chain;
  const uint8_t* link;     // [n] 1: node i-1 is node i's predecessor on a path of the cover (never set when i % 16 == 0)
  const int32_t* es_ptr;   // CSR: node i sums the coupling blocks es[es_ptr[i] .. es_ptr[i+1])
  const int32_t* es_slot;  // [m] -1, or 2 * (index into es) + (1 if the factor's node a is the later node of the pair)
  double* es;              // [<= m][16] J_i^T J_{i-1} of the chain factors (written at every linearisation)
  double* En;              // [n][16] summed coupling block of node i with node i-1
};

__device__ __forceinline__ void all_sync(const SolverDev& P, cg::grid_group& grid) {
  if (P.use_cluster) cg::this_cluster().sync(); else grid.sync();
}

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_t(double v) { return warp_sum_d(v); }

// grid-wide sums of K values.  The warp level runs in T (cheap in fp32), the 8 warp sums, the CTA partials and the final
// sum are fp64 (a handful of instructions on warp 0 only).
template <int K, typename T>
__device__ void grid_reduce_sum(T (&vin)[K], double (&v)[K], const SolverDev& P, int parity, double* sh, cg::grid_group& grid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = GS_THREADS / 32, G = gridDim.x;
  double* pbuf = P.partial + (size_t)parity * 4 * G;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const T w = warp_sum_t(vin[k]);
    if (lane == 0) sh[k * 32 + warp] = (double)w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double xv = (lane < nw) ? sh[k * 32 + lane] : 0.0;
      xv = warp_sum_d(xv);
      if (lane == 0) __stcg(pbuf + k * G + blockIdx.x, xv);
    }
  }
  all_sync(P, grid);
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double xv = 0.0;
      for (int i = lane; i < G; i += 32) xv += __ldcg(pbuf + k * G + i);
      xv = warp_sum_d(xv);
      if (lane == 0) sh[k] = xv;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = sh[k];
  __syncthreads();
}

__device__ double grid_reduce_max(double vmax, const SolverDev& P, int parity, double* sh, cg::grid_group& grid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  __syncwarp();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane == 0) sh[warp] = vmax;
  __syncthreads();
  double* pbuf = P.partial + (size_t)parity * 4 * G;
  if (warp == 0) {
    double xv = (lane < GS_THREADS / 32) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xv = fmax(xv, __shfl_xor_sync(0xffffffffu, xv, o));
    if (lane == 0) __stcg(pbuf + blockIdx.x, xv);
  }
  all_sync(P, grid);
  if (warp == 0) {
    double xv = 0.0;
    for (int i = lane; i < G; i += 32) xv = fmax(xv, __ldcg(pbuf + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xv = fmax(xv, __shfl_xor_sync(0xffffffffu, xv, o));
    if (lane == 0) sh[0] = xv;
  }
  __syncthreads();
  const double r = sh[0];
  __syncthreads();
  return r;
}

// Jacobian store accessor: element i (0..31: Ja row-major then Jb) of the factor with CTA-local index `li`
template <typename T>
struct JStore {
  T* base; int stride; int off;
  __device__ __forceinline__ T& at(int i, int li) const { return base[(size_t)i * stride + off + li]; }
};

// clock read that cannot be scheduled before `v` has been computed (phase attribution only)
__device__ __forceinline__ long long clock_after(float v) {
  long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) : "f"(v) : "memory"); return c;
}
__device__ __forceinline__ long long clock_after(double v) {
  long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) : "d"(v) : "memory"); return c;
}

// vector loads / stores of a node's 4-vector (one 16-byte access in fp32, two in fp64), L2-coherent
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  __stcg(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
  __stcg(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
  __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(v[2], v[3]));
}

// 4x4 SPD inverse by Gauss-Jordan (no pivoting: the LM term keeps the diagonal positive)
template <typename T>
__device__ void inv4(const T* __restrict__ M, T* __restrict__ out) {
  T a[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[i][j] = M[i * 4 + j]; a[i][4 + j] = (i == j) ? T(1) : T(0); }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const T piv = T(1) / a[c][c];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[c][j] *= piv;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i == c) continue;
      const T f = a[i][c];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[i][j] -= f * a[c][j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = a[i][4 + j];
}

// ---- chain preconditioner -----------------------------------------------------------------------------------------
// Block-Jacobi sees only a node's own 4x4 block, and a pose graph is dominated by long odometry chains: C5 needs 1663
// PCG iterations that way.  The host covers the graph with vertex-disjoint paths (heaviest factors first) and numbers the
// nodes along them, so a path is a run of consecutive node ids = consecutive threads.  The preconditioner is the block-
// TRIDIAGONAL part of J^T J + lam D along those paths, cut every 16 nodes so that a segment lives in one half-warp:
//   M = sum over chain factors (full 8x8 contribution) + sum over the other factors (their two diagonal blocks) + lam D,
// a sum of PSD terms plus a positive diagonal, hence SPD.  Factorisation M = (I + L) S (I + L)^T by a 16-step sweep over
// the half-warp (once per PCG solve); application z = M^-1 r = 15 forward + 15 backward steps of 4-vector shuffles.
// Same CPU emulation as the kernel: 1663 -> 385 PCG iterations on C5.
template <typename T>
__device__ __forceinline__ void mat4_mul(const T* A, const T* B, T* C) {        // C = A B
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T a = T(0);
#pragma unroll
      for (int k = 0; k < 4; ++k) a += A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = a;
    }
}

// z = M^-1 r for the segment this half-warp owns.  Si = S_i^-1, L = L_i (zero when the node has no predecessor link).
// Warp-collective: every lane of the warp must call it.
template <typename T>
__device__ __forceinline__ void chain_apply(const T (&Si)[16], const T (&L)[16], bool link, int sl, const T (&r)[4], T (&z)[4]) {
  // Branch-free on purpose (no divergent code between two shuffles): the active lane of a step is selected by a 0/1
  // multiplier; L is zero on lanes without a predecessor link.
  __syncwarp();
  T y[4] = {r[0], r[1], r[2], r[3]};
#pragma unroll
  for (int s = 1; s < 16; ++s) {                       // forward: y_i = r_i - L_i y_{i-1}
    T yp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) yp[k] = __shfl_up_sync(0xffffffffu, y[k], 1);
    const T m = (sl == s && link) ? T(1) : T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      y[i] -= m * (L[i * 4] * yp[0] + L[i * 4 + 1] * yp[1] + L[i * 4 + 2] * yp[2] + L[i * 4 + 3] * yp[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = Si[i * 4] * y[0] + Si[i * 4 + 1] * y[1] + Si[i * 4 + 2] * y[2] + Si[i * 4 + 3] * y[3];
#pragma unroll
  for (int s = 14; s >= 0; --s) {                      // backward: z_i = S_i^-1 y_i - L_{i+1}^T z_{i+1}
    const T m = (sl == s + 1 && link) ? T(1) : T(0);
    T u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = m * (L[j] * z[0] + L[4 + j] * z[1] + L[8 + j] * z[2] + L[12 + j] * z[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __shfl_down_sync(0xffffffffu, u[k], 1);
    const T mz = (sl == s) ? T(1) : T(0);              // (lane 15 receives lane 16's u, which is zero: sl = 0 there)
#pragma unroll
    for (int k = 0; k < 4; ++k) z[k] -= mz * u[k];
  }
}

// Linearise every factor of this CTA at `xp` (fp64): robustified Jacobians -> J store (as T), gradient contributions
// J^T r -> gs slots, diagonal-block contributions J^T J -> hs slots, chain couplings -> es.  Returns this thread's share
// of the cost.
template <typename T>
__device__ double factor_linearize(const SolverDev& P, const double* __restrict__ xp, const JStore<T>& J) {
  double cost = 0.0;
  const int f0 = blockIdx.x * P.fpc, f1 = min(P.m, f0 + P.fpc);
  for (int f = f0 + threadIdx.x; f < f1; f += GS_THREADS) {
    const int li = f - f0;
    const int a = P.ia[f], b = P.ib[f];
    double pa[4], pb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { pa[i] = __ldcg(xp + 4 * a + i); pb[i] = __ldcg(xp + 4 * b + i); }
    double r[4], Ja[16], Jb[16];
    linearize_factor(P.ftype[f], pa, pb, P.payload + (size_t)f * OSB_PAYLOAD_LEN, r, Ja, Jb);
    const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
    double w = 1.0;
    if (P.huber[f] && s > 1.0) {        // ceres::HuberLoss(1.0): rho(s) = 2 sqrt(s) - 1, sqrt(rho') = s^-1/4
      cost += 0.5 * (2.0 * sqrt(s) - 1.0);
      w = 1.0 / sqrt(sqrt(s));
    } else {
      cost += 0.5 * s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] *= w;
#pragma unroll
    for (int i = 0; i < 16; ++i) { Ja[i] *= w; Jb[i] *= w; J.at(i, li) = (T)Ja[i]; J.at(16 + i, li) = (T)Jb[i]; }
    if (P.use_chain) {
      const int es = P.es_slot[f];
      if (es >= 0) {                      // E = J_later^T J_earlier (rows: the later node of the pair)
        const double* Jl = (es & 1) ? Ja : Jb;
        const double* Je = (es & 1) ? Jb : Ja;
        double* dst = P.es + 16 * (size_t)(es >> 1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) t += Jl[i * 4 + j] * Je[i * 4 + k];
            __stcg(dst + j * 4 + k, t);
          }
      }
    }
    double* ga = P.gs + 4 * (size_t)P.slot_a[f];
    double* gb = P.gs + 4 * (size_t)P.slot_b[f];
    double* ha = P.hs + 16 * (size_t)P.slot_a[f];
    double* hb = P.hs + 16 * (size_t)P.slot_b[f];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) { sa += Ja[i * 4 + j] * r[i]; sb += Jb[i * 4 + j] * r[i]; }
      __stcg(ga + j, sa); __stcg(gb + j, sb);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { ta += Ja[i * 4 + j] * Ja[i * 4 + k]; tb += Jb[i * 4 + j] * Jb[i * 4 + k]; }
        __stcg(ha + j * 4 + k, ta); __stcg(hb + j * 4 + k, tb);
      }
    }
  }
  return cost;
}

// cost at the trial point `xn` (fp64) and this thread's share of |J_cur delta|^2 (model decrease, in T)
template <typename T>
__device__ void factor_trial(const SolverDev& P, const double* __restrict__ xn, const JStore<T>& J, double& cost, double& jd) {
  const int f0 = blockIdx.x * P.fpc, f1 = min(P.m, f0 + P.fpc);
  const T* delta = static_cast<const T*>(P.delta);
  T jdt = T(0);
  for (int f = f0 + threadIdx.x; f < f1; f += GS_THREADS) {
    const int li = f - f0;
    const int a = P.ia[f], b = P.ib[f];
    double pa[4], pb[4];
    T da[4], db[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { pa[i] = __ldcg(xn + 4 * a + i); pb[i] = __ldcg(xn + 4 * b + i); }
    ld4(delta + 4 * a, da); ld4(delta + 4 * b, db);
    const double s = factor_sqnorm(P.ftype[f], pa, pb, P.payload + (size_t)f * OSB_PAYLOAD_LEN);
    cost += (P.huber[f] && s > 1.0) ? 0.5 * (2.0 * sqrt(s) - 1.0) : 0.5 * s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      T t = T(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) t += J.at(i * 4 + j, li) * da[j] + J.at(16 + i * 4 + j, li) * db[j];
      jdt += t * t;
    }
  }
  jd += (double)jdt;
}

template <typename T>
__global__ void __launch_bounds__(GS_THREADS, 1)
graph_solve_kernel(SolverDev P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem_j = reinterpret_cast<T*>(smem_raw);
  __shared__ double sh[4 * 32];
  const int T_ = gridDim.x * GS_THREADS;
  const int gtid = blockIdx.x * GS_THREADS + threadIdx.x;
  JStore<T> J;
  if (P.j_in_smem) { J.base = smem_j; J.stride = P.fpc; J.off = 0; }
  else { J.base = static_cast<T*>(P.Jg); J.stride = P.m; J.off = blockIdx.x * P.fpc; }
  T* Ls = smem_j + (size_t)32 * P.fpc;          // [16][GS_THREADS]: L_i of the chain preconditioner (use_chain only)
  T* const Pp = static_cast<T*>(P.p); T* const Pz = static_cast<T*>(P.z); T* const Pres = static_cast<T*>(P.res);
  T* const PAp = static_cast<T*>(P.Ap); T* const Pdelta = static_cast<T*>(P.delta); T* const PMinv = static_cast<T*>(P.Minv);
  T* const Pcs = static_cast<T*>(P.cs);
  int parity = 0;
  unsigned long long t0 = 0;
  const long long k0 = clock64();
  if (gtid == 0) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < 8 + GS_MAX_CLUSTER * (GS_THREADS / 32); ++i) P.dbg[i] = 0;
  }

  int cur = 0;
  double radius = P.opt.initial_trust_radius;
  double decrease = 2.0;
  int iters = 0, pcg_total = 0, termination = 3;

  // ---- initial linearisation ----
  double v1[1], w1[1] = {factor_linearize(P, P.x[0], J)};
  grid_reduce_sum<1>(w1, v1, P, parity, sh, grid); parity ^= 1;
  double cost = v1[0];
  const double initial_cost = cost;
  bool need_gradient = true;
  const int f0 = blockIdx.x * P.fpc, f1 = min(P.m, f0 + P.fpc);
  // static per-thread data of the fast CG path
  const bool fast = (P.fpc <= GS_KF * GS_THREADS) && (P.n <= T_);
  int fa[GS_KF], fb[GS_KF], fsa[GS_KF], fsb[GS_KF];
  bool fvalid[GS_KF];
#pragma unroll
  for (int k = 0; k < GS_KF; ++k) {
    fvalid[k] = false; fa[k] = fb[k] = fsa[k] = fsb[k] = 0;
    const int f = f0 + threadIdx.x + k * GS_THREADS;
    if (fast && f < f1) { fvalid[k] = true; fa[k] = P.ia[f]; fb[k] = P.ib[f]; fsa[k] = P.slot_a[f]; fsb[k] = P.slot_b[f]; }
  }
  const bool is_node = fast && gtid < P.n && !P.fixed[gtid];
  const int ns0 = is_node ? P.node_ptr[gtid] : 0, ns1 = is_node ? P.node_ptr[gtid + 1] : 0;
  const bool chain = fast && P.use_chain;
  const int sl = threadIdx.x & 15;

  while (iters < P.opt.max_iterations) {
    if (need_gradient) {
      // ---- node phase G (fp64): gather gradient and diagonal blocks from the slots, LM diagonal ----
      double vmax = 0.0;
      for (int n = gtid; n < P.n; n += T_) {
        double gn[4] = {0.0, 0.0, 0.0, 0.0};
        double Hn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) Hn[i] = 0.0;
        if (!P.fixed[n]) {
          const int s0 = P.node_ptr[n], s1 = P.node_ptr[n + 1];
#pragma unroll 2
          for (int sidx = s0; sidx < s1; ++sidx) {
#pragma unroll
            for (int i = 0; i < 4; ++i) gn[i] += __ldcg(P.gs + 4 * (size_t)sidx + i);
#pragma unroll
            for (int i = 0; i < 16; ++i) Hn[i] += __ldcg(P.hs + 16 * (size_t)sidx + i);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          P.g[4 * n + i] = gn[i];
          P.D[4 * n + i] = fmin(fmax(Hn[i * 5], 1e-6), 1e32);
          vmax = fmax(vmax, fabs(gn[i]));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) P.Hnn[16 * n + i] = Hn[i];
        if (P.use_chain) {
          double En[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) En[i] = 0.0;
          for (int e = P.es_ptr[n]; e < P.es_ptr[n + 1]; ++e)
#pragma unroll
            for (int i = 0; i < 16; ++i) En[i] += __ldcg(P.es + 16 * (size_t)e + i);
#pragma unroll
          for (int i = 0; i < 16; ++i) P.En[16 * n + i] = En[i];
        }
      }
      const double gmax = grid_reduce_max(vmax, P, parity, sh, grid); parity ^= 1;
      need_gradient = false;
      if (gmax <= P.opt.gradient_tolerance) { termination = 1; break; }
    }

    // ---- PCG init (node phase): preconditioner, res = -g, z = M^-1 res, p = 0, delta = 0 ----
    // Fast path (every node has its own thread, <= GS_KF factors per thread): the node's preconditioner blocks, D, res,
    // z, p, delta stay in REGISTERS for the whole CG solve and the factor's indices are preloaded, so a CG iteration
    // touches global memory only for what crosses threads: z/p gathers by the factor threads and the contribution slots.
    const double lam = 1.0 / radius;
    T v2[2] = {T(0), T(0)};
    T Mi[16], Dl[4] = {T(0), T(0), T(0), T(0)}, rn[4] = {T(0), T(0), T(0), T(0)}, zn[4] = {T(0), T(0), T(0), T(0)};
    T pn[4] = {T(0), T(0), T(0), T(0)}, dn[4] = {T(0), T(0), T(0), T(0)};     // Dl = lam * D
    bool lk = false;
    if (chain) {
      // factorise the segment's block-tridiagonal M = (I + L) S (I + L)^T: 16 steps over the half-warp, Mi := S_i^-1
      T M[16], E[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { M[i] = (i % 5 == 0) ? T(1) : T(0); E[i] = T(0); Mi[i] = T(0); }
      if (is_node) {
        double Md[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) Md[i] = P.Hnn[16 * gtid + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double ld = lam * P.D[4 * gtid + i]; Dl[i] = (T)ld; Md[i * 5] += ld; }
#pragma unroll
        for (int i = 0; i < 16; ++i) M[i] = (T)Md[i];
        lk = P.link[gtid] != 0;
        if (lk) {
#pragma unroll
          for (int i = 0; i < 16; ++i) E[i] = (T)P.En[16 * gtid + i];
        }
      }
      T Lr[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) Lr[i] = T(0);
      for (int s = 0; s < 16; ++s) {
        T Sp[16];
        __syncwarp();                                              // converge after the previous step's divergent block
#pragma unroll
        for (int i = 0; i < 16; ++i) Sp[i] = __shfl_up_sync(0xffffffffu, Mi[i], 1);
        if (sl == s) {
          if (lk) {
            mat4_mul(E, Sp, Lr);                                   // L_i = E_i S_{i-1}^-1
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) {                        // S_i = M_i - L_i E_i^T
                T a = T(0);
#pragma unroll
                for (int k = 0; k < 4; ++k) a += Lr[i * 4 + k] * E[j * 4 + k];
                M[i * 4 + j] -= a;
              }
          }
          inv4(M, Mi);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) Ls[i * GS_THREADS + threadIdx.x] = Lr[i];
      if (is_node) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rn[i] = (T)(-P.g[4 * gtid + i]);
      }
      chain_apply(Mi, Lr, lk, sl, rn, zn);
      if (gtid < P.n) {
        const T zero4[4] = {T(0), T(0), T(0), T(0)};
        st4(Pres + 4 * gtid, rn); st4(Pz + 4 * gtid, zn); st4(Pp + 4 * gtid, zero4); st4(Pdelta + 4 * gtid, zero4);
      }
      if (is_node) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v2[0] += rn[i] * zn[i]; v2[1] += rn[i] * rn[i]; }
      }
    } else {
      for (int n = gtid; n < P.n; n += T_) {
        T rl[4] = {T(0), T(0), T(0), T(0)}, zl[4] = {T(0), T(0), T(0), T(0)};
        if (!P.fixed[n]) {
          double Md[16];
          T M[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) Md[i] = P.Hnn[16 * n + i];
#pragma unroll
          for (int i = 0; i < 4; ++i) { const double ld = lam * P.D[4 * n + i]; Dl[i] = (T)ld; Md[i * 5] += ld; }
#pragma unroll
          for (int i = 0; i < 16; ++i) M[i] = (T)Md[i];
          inv4(M, Mi);
#pragma unroll
          for (int i = 0; i < 16; ++i) PMinv[16 * n + i] = Mi[i];
#pragma unroll
          for (int i = 0; i < 4; ++i) rl[i] = (T)(-P.g[4 * n + i]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) zl[i] += Mi[i * 4 + j] * rl[j];
        }
        const T zero4[4] = {T(0), T(0), T(0), T(0)};
        st4(Pres + 4 * n, rl); st4(Pz + 4 * n, zl); st4(Pp + 4 * n, zero4); st4(Pdelta + 4 * n, zero4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v2[0] += rl[i] * zl[i]; v2[1] += rl[i] * rl[i];
          rn[i] = rl[i]; zn[i] = zl[i];                       // (fast path: the only iteration of this loop)
        }
      }
    }
    double r2[2];
    grid_reduce_sum<2>(v2, r2, P, parity, sh, grid); parity ^= 1;
    double rz = r2[0];
    const double rr0 = r2[1];
    T beta = T(0);
    int it = 0;
    // ---- PCG iterations: 3 barriers each ----
    while (it < P.opt.max_pcg_iterations && rr0 > 0.0) {
      long long c0 = clock64();
      // factor phase: p_a = z_a + beta p_a (on the fly), t = Ja p_a + Jb p_b, contributions Ja^T t, Jb^T t
      if (fast) {
        // all gathers of this thread's (up to 3) factors are issued before any arithmetic: one L2 round trip
        T zq[GS_KF][2][4], pq[GS_KF][2][4];
#pragma unroll
        for (int k = 0; k < GS_KF; ++k) {
          ld4(Pz + 4 * fa[k], zq[k][0]); ld4(Pz + 4 * fb[k], zq[k][1]);
          ld4(Pp + 4 * fa[k], pq[k][0]); ld4(Pp + 4 * fb[k], pq[k][1]);
        }
#pragma unroll
        for (int k = 0; k < GS_KF; ++k) {
          if (!fvalid[k]) continue;
          const int li = threadIdx.x + k * GS_THREADS;
          T pa[4], pb[4], t[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { pa[i] = zq[k][0][i] + beta * pq[k][0][i]; pb[i] = zq[k][1][i] + beta * pq[k][1][i]; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += J.at(i * 4 + j, li) * pa[j] + J.at(16 + i * 4 + j, li) * pb[j];
            t[i] = acc;
          }
          T ca[4], cb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            T sa = T(0), sb = T(0);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sa += J.at(i * 4 + j, li) * t[i]; sb += J.at(16 + i * 4 + j, li) * t[i]; }
            ca[j] = sa; cb[j] = sb;
          }
          st4(Pcs + 4 * (size_t)fsa[k], ca);
          st4(Pcs + 4 * (size_t)fsb[k], cb);
        }
      } else {
        for (int f = f0 + threadIdx.x; f < f1; f += GS_THREADS) {
          const int li = f - f0;
          const int a = P.ia[f], b = P.ib[f];
          T za[4], zb[4], qa[4], qb[4], pa[4], pb[4], t[4];
          ld4(Pz + 4 * a, za); ld4(Pz + 4 * b, zb); ld4(Pp + 4 * a, qa); ld4(Pp + 4 * b, qb);
#pragma unroll
          for (int i = 0; i < 4; ++i) { pa[i] = za[i] + beta * qa[i]; pb[i] = zb[i] + beta * qb[i]; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += J.at(i * 4 + j, li) * pa[j] + J.at(16 + i * 4 + j, li) * pb[j];
            t[i] = acc;
          }
          T ca[4], cb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            T sa = T(0), sb = T(0);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sa += J.at(i * 4 + j, li) * t[i]; sb += J.at(16 + i * 4 + j, li) * t[i]; }
            ca[j] = sa; cb[j] = sb;
          }
          st4(Pcs + 4 * (size_t)P.slot_a[f], ca);
          st4(Pcs + 4 * (size_t)P.slot_b[f], cb);
        }
      }
      long long c1 = clock64();
      all_sync(P, grid);
      long long c2 = clock64();
      // node phase 1: p = z + beta p, Ap = sum of the node's slots + lam D p, partial p.Ap
      T v1b[1] = {T(0)};
      T apn[4] = {T(0), T(0), T(0), T(0)};
      if (fast) {
        if (is_node) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { pn[i] = zn[i] + beta * pn[i]; apn[i] = Dl[i] * pn[i]; }
          st4(Pp + 4 * gtid, pn);
          for (int s0 = ns0; s0 < ns1; s0 += 16) {         // 16 (fp32) / 32 (fp64) independent 16-byte loads in flight
            T c[16][4];                                    // per batch: one L2 round trip covers every node of degree <= 16
#pragma unroll
            for (int k = 0; k < 16; ++k) ld4(Pcs + 4 * (size_t)min(s0 + k, ns1 - 1), c[k]);
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (s0 + k < ns1) { apn[0] += c[k][0]; apn[1] += c[k][1]; apn[2] += c[k][2]; apn[3] += c[k][3]; }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) v1b[0] += pn[i] * apn[i];
        }
      } else {
        for (int n = gtid; n < P.n; n += T_) {
          if (P.fixed[n]) continue;
          T zl[4], ql[4], pl[4], ap[4];
          ld4(Pz + 4 * n, zl); ld4(Pp + 4 * n, ql);
#pragma unroll
          for (int i = 0; i < 4; ++i) { pl[i] = zl[i] + beta * ql[i]; ap[i] = (T)(lam * P.D[4 * n + i]) * pl[i]; }
          const int s0 = P.node_ptr[n], s1 = P.node_ptr[n + 1];
#pragma unroll 4
          for (int sidx = s0; sidx < s1; ++sidx) {
            T c[4];
            ld4(Pcs + 4 * (size_t)sidx, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) ap[i] += c[i];
          }
          st4(Pp + 4 * n, pl); st4(PAp + 4 * n, ap);
#pragma unroll
          for (int i = 0; i < 4; ++i) v1b[0] += pl[i] * ap[i];
        }
      }
      long long c3 = clock_after(v1b[0]);
      double r1[1];
      grid_reduce_sum<1>(v1b, r1, P, parity, sh, grid); parity ^= 1;
      long long c4 = clock64();
      const double pAp = r1[0];
      if (!(pAp > 0.0)) break;
      const T alpha = (T)(rz / pAp);
      // node phase 2: delta += alpha p, res -= alpha Ap, z = M^-1 res; partial rz_new, rr
      T v22[2] = {T(0), T(0)};
      if (fast) {
        // (no is_node guard: pn / apn are zero on the other lanes, and a divergent branch here would send the warp
        //  through the slow collective-shuffle path of chain_apply)
#pragma unroll
        for (int i = 0; i < 4; ++i) { dn[i] += alpha * pn[i]; rn[i] -= alpha * apn[i]; }
        if (chain) {                                          // warp-collective: every lane takes part in the sweeps
          T Lr[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) Lr[i] = Ls[i * GS_THREADS + threadIdx.x];
          const long long q0 = clock_after(rn[0] + Lr[15]);
          chain_apply(Mi, Lr, lk, sl, rn, zn);
          const long long q1 = clock_after(zn[0] + zn[3]);
          if ((threadIdx.x & 31) == 0) P.dbg[8 + blockIdx.x * (GS_THREADS / 32) + (threadIdx.x >> 5)] += q1 - q0;
        } else if (is_node) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += Mi[i * 4 + j] * rn[j];
            zn[i] = acc;
          }
        }
        if (is_node) {
          st4(Pz + 4 * gtid, zn);
#pragma unroll
          for (int i = 0; i < 4; ++i) { v22[0] += rn[i] * zn[i]; v22[1] += rn[i] * rn[i]; }
        }
      } else {
        for (int n = gtid; n < P.n; n += T_) {
          if (P.fixed[n]) continue;
          T dl[4], pl[4], rl[4], al[4], zl[4] = {T(0), T(0), T(0), T(0)};
          ld4(Pdelta + 4 * n, dl); ld4(Pp + 4 * n, pl); ld4(Pres + 4 * n, rl); ld4(PAp + 4 * n, al);
#pragma unroll
          for (int i = 0; i < 4; ++i) { dl[i] += alpha * pl[i]; rl[i] -= alpha * al[i]; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) zl[i] += PMinv[16 * n + i * 4 + j] * rl[j];
          st4(Pdelta + 4 * n, dl); st4(Pres + 4 * n, rl); st4(Pz + 4 * n, zl);
#pragma unroll
          for (int i = 0; i < 4; ++i) { v22[0] += rl[i] * zl[i]; v22[1] += rl[i] * rl[i]; }
        }
      }
      long long c5 = clock_after(v22[0] + v22[1]);
      double r22[2];
      grid_reduce_sum<2>(v22, r22, P, parity, sh, grid); parity ^= 1;
      if (gtid == 0) {
        const long long c6 = clock64();
        P.dbg[0] += c1 - c0; P.dbg[1] += c2 - c1; P.dbg[2] += c3 - c2; P.dbg[3] += c4 - c3; P.dbg[4] += c5 - c4;
        P.dbg[5] += c6 - c5; P.dbg[6] += 1;
      }
      ++it;
      beta = (T)(r22[0] / rz);
      rz = r22[0];
      if (r22[1] <= P.opt.pcg_tolerance * P.opt.pcg_tolerance * rr0) break;
    }
    if (fast && is_node) st4(Pdelta + 4 * gtid, dn);
    pcg_total += it;
    ++iters;

    // ---- trial point (fp64): x_new = x + delta; g.delta, |delta|^2, |x|^2 ----
    double* xc = P.x[cur];
    double* xn = P.x[cur ^ 1];
    double v4[3] = {0.0, 0.0, 0.0}, r4[3];
    for (int n = gtid; n < P.n; n += T_) {
      T dl[4];
      ld4(Pdelta + 4 * n, dl);
      if (fast && n == gtid && is_node) { dl[0] = dn[0]; dl[1] = dn[1]; dl[2] = dn[2]; dl[3] = dn[3]; }   // (own store)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double d = P.fixed[n] ? 0.0 : (double)dl[i];
        const double xv = xc[4 * n + i];
        __stcg(xn + 4 * n + i, xv + d);
        v4[0] += P.g[4 * n + i] * d; v4[1] += d * d;
        if (!P.fixed[n]) v4[2] += xv * xv;
      }
      if (P.fixed[n]) { const T zero4[4] = {T(0), T(0), T(0), T(0)}; st4(Pdelta + 4 * n, zero4); }
    }
    grid_reduce_sum<3>(v4, r4, P, parity, sh, grid); parity ^= 1;
    // ---- evaluate trial, model decrease, elapsed time ----
    double v5[3] = {0.0, 0.0, 0.0}, r5[3];
    factor_trial(P, xn, J, v5[0], v5[1]);
    if (gtid == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      v5[2] = (double)(t1 - t0) * 1e-9;
    }
    grid_reduce_sum<3>(v5, r5, P, parity, sh, grid); parity ^= 1;
    const double new_cost = r5[0];
    const double model = -r4[0] - 0.5 * r5[1];
    const double elapsed = r5[2];
    const double rho = (model > 0.0) ? (cost - new_cost) / model : -1.0;
    const bool finite = isfinite(new_cost);
    const bool small_step = sqrt(r4[1]) <= P.opt.parameter_tolerance * (sqrt(r4[2]) + P.opt.parameter_tolerance);
    if (finite && rho > 1e-3) {
      // accept (Ceres LevenbergMarquardtStrategy::StepAccepted)
      const double tmp = 2.0 * rho - 1.0;
      radius = fmin(radius / fmax(1.0 / 3.0, 1.0 - tmp * tmp * tmp), 1e16);
      decrease = 2.0;
      const double dcost = cost - new_cost;
      const double old_cost = cost;
      cur ^= 1;
      cost = new_cost;
      if (fabs(dcost) <= P.opt.function_tolerance * old_cost) { termination = 0; break; }
      if (small_step) { termination = 2; break; }
      if (P.opt.max_time_s > 0.0 && elapsed > P.opt.max_time_s) { termination = 4; break; }
      // re-linearise at the accepted point (the trial pass kept the old Jacobians for the model term)
      double v1c[1], w1c[1] = {factor_linearize(P, P.x[cur], J)};
      grid_reduce_sum<1>(w1c, v1c, P, parity, sh, grid); parity ^= 1;
      need_gradient = true;
    } else {
      radius /= decrease;
      decrease *= 2.0;
      if (radius < 1e-32 || !isfinite(radius)) { termination = 5; break; }
      if (small_step) { termination = 2; break; }
      if (P.opt.max_time_s > 0.0 && elapsed > P.opt.max_time_s) { termination = 4; break; }
    }
  }

  // ---- write back ----
  const double* xf = P.x[cur];
  for (int i = gtid; i < 4 * P.n; i += T_) P.poses_out[i] = __ldcg(xf + i);
  if (gtid == 0) {
    P.summary->initial_cost = initial_cost;
    P.summary->final_cost = cost;
    P.summary->iterations = iters;
    P.summary->pcg_iterations = pcg_total;
    P.summary->termination = termination;
    P.dbg[7] = clock64() - k0;
  }
}

__global__ void graph_linearize_kernel(int m, const double* __restrict__ poses, const int32_t* __restrict__ ftype,
                                       const int32_t* __restrict__ ia, const int32_t* __restrict__ ib,
                                       const double* __restrict__ payload, double* __restrict__ r,
                                       double* __restrict__ Ja, double* __restrict__ Jb) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= m) return;
  double rr[4], A[16], B[16];
  linearize_factor(ftype[f], poses + 4 * ia[f], poses + 4 * ib[f], payload + (size_t)f * OSB_PAYLOAD_LEN, rr, A, B);
  for (int i = 0; i < 4; ++i) r[(size_t)f * 4 + i] = rr[i];
  for (int i = 0; i < 16; ++i) { Ja[(size_t)f * 16 + i] = A[i]; Jb[(size_t)f * 16 + i] = B[i]; }
}

}  // namespace osb

using namespace osb;

struct osb_solver {
  int max_nodes = 0, max_factors = 0;
  bool cluster_ok = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::mutex mu;
  // device
  uint8_t *d_fixed = nullptr, *d_huber = nullptr;
  int32_t *d_type = nullptr, *d_ia = nullptr, *d_ib = nullptr, *d_ptr = nullptr, *d_slot_a = nullptr, *d_slot_b = nullptr;
  double *d_payload = nullptr, *d_x0 = nullptr, *d_x1 = nullptr, *d_Jg = nullptr, *d_lin = nullptr;
  double *d_nodevec = nullptr;   // g, D, p, z, res, Ap, delta (7 x 4n) + Hnn, Minv (2 x 16n)
  double *d_cs = nullptr, *d_gs = nullptr, *d_hs = nullptr, *d_partial = nullptr, *d_out = nullptr;
  osb_solve_summary* d_summary = nullptr;
  long long* d_dbg = nullptr;
  uint8_t* d_link = nullptr;
  // pinned staging for what crosses PCIe on EVERY solve (poses in, poses + summary out): a copy to / from pageable memory
  // blocks inside the driver until the stream reaches it -- for the results that is the whole solve, and other host threads
  // (the keyframe front-end of the same process) could not launch meanwhile (measured: 640 -> 57 keyframes/s beside a
  // back-to-back solver thread).  Pinned copies are asynchronous; the one wait is a cudaStreamSynchronize.
  double* h_x = nullptr;
  osb_solve_summary* h_summary = nullptr;
  int device = 0;
  int32_t *d_es_ptr = nullptr, *d_es_slot = nullptr;
  double *d_es = nullptr, *d_En = nullptr;
  int last_grid = 0, last_cluster = 0, last_jsmem = 0, last_chain = 0, last_f32 = 0;
  // resident graph (osb_solver_graph_*): host mirrors of what is already in device memory
  std::vector<double> g_poses, g_payload;
  std::vector<uint8_t> g_fixed, g_huber;
  std::vector<int32_t> g_type, g_ia, g_ib;
  size_t g_uploaded = 0;            // factors whose type / huber / payload are already on the device
  bool g_static_valid = false;      // false after a one-shot osb_solver_solve overwrote the device factor arrays
  // topology cache of the resident graph: the path cover, the internal numbering and the CSR / slot tables depend only on
  // (fixed, type, ia, ib, payload weights), not on the poses -- a window that is re-solved without new frames (or after
  // set_poses) re-uses them on the host AND on the device (0.5 ms of host work per solve otherwise)
  unsigned long long g_topo_version = 1, cached_topo = 0;
  std::vector<int32_t> cached_order;
  int cached_nres = 0;
};

extern "C" void osb_solve_default_options(osb_solve_options* o) {
  if (!o) return;
  o->max_iterations = 1000;         // swarm_localization_solver.cpp:1697
  o->max_pcg_iterations = 500;
  o->max_time_s = 0.0;
  o->function_tolerance = 1e-6;     // Ceres defaults
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->pcg_tolerance = 1e-2;
  o->preconditioner = OSB_PRECOND_AUTO;
  o->inner_precision = OSB_INNER_AUTO;
  o->initial_trust_radius = 1e4;
}

extern "C" osb_status osb_solver_create(osb_solver** out, int max_nodes, int max_factors) {
  OSB_REQUIRE(out != nullptr && max_nodes > 0 && max_factors > 0, "bad sizes");
  osb_status s = require_device();
  if (s != OSB_OK) return s;
  osb_solver* h = new osb_solver();
  h->max_nodes = max_nodes; h->max_factors = max_factors;
  const size_t n = max_nodes, m = max_factors;
  OSB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  OSB_CUDA(cudaEventCreate(&h->ev0));
  OSB_CUDA(cudaEventCreate(&h->ev1));
  OSB_CUDA(cudaFuncSetAttribute(graph_solve_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_DYN_MAX));
  OSB_CUDA(cudaFuncSetAttribute(graph_solve_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_DYN_MAX));
  h->cluster_ok = cudaFuncSetAttribute(graph_solve_kernel<float>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                  cudaFuncSetAttribute(graph_solve_kernel<double>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  cudaGetLastError();
  OSB_CUDA(cudaMalloc(&h->d_fixed, n));
  OSB_CUDA(cudaMalloc(&h->d_huber, m));
  OSB_CUDA(cudaMalloc(&h->d_type, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_ia, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_ib, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_slot_a, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_slot_b, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_ptr, (n + 1) * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_payload, m * OSB_PAYLOAD_LEN * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_x0, 4 * n * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_x1, 4 * n * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_Jg, 32 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_lin, 36 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_nodevec, (7 * 4 + 2 * 16) * n * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_cs, 8 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_gs, 8 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_hs, 32 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_partial, 2 * 4 * (size_t)num_sms() * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_out, 4 * n * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_summary, sizeof(osb_solve_summary)));
  OSB_CUDA(cudaMalloc(&h->d_dbg, (8 + 128) * sizeof(long long)));
  OSB_CUDA(cudaMemset(h->d_dbg, 0, (8 + 128) * sizeof(long long)));
  OSB_CUDA(cudaMalloc(&h->d_link, n));
  OSB_CUDA(cudaMalloc(&h->d_es_ptr, (n + 1) * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_es_slot, m * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&h->d_es, 16 * m * sizeof(double)));
  OSB_CUDA(cudaMalloc(&h->d_En, 16 * n * sizeof(double)));
  OSB_CUDA(cudaHostAlloc((void**)&h->h_x, 4 * n * sizeof(double), cudaHostAllocDefault));
  OSB_CUDA(cudaHostAlloc((void**)&h->h_summary, sizeof(osb_solve_summary), cudaHostAllocDefault));
  h->device = current_device();
  *out = h;
  return OSB_OK;
}

extern "C" osb_status osb_solver_destroy(osb_solver* h) {
  if (!h) return OSB_OK;
  cudaFree(h->d_fixed); cudaFree(h->d_huber); cudaFree(h->d_type); cudaFree(h->d_ia); cudaFree(h->d_ib);
  cudaFree(h->d_slot_a); cudaFree(h->d_slot_b); cudaFree(h->d_ptr); cudaFree(h->d_payload); cudaFree(h->d_x0);
  cudaFree(h->d_x1); cudaFree(h->d_Jg); cudaFree(h->d_lin); cudaFree(h->d_nodevec); cudaFree(h->d_cs); cudaFree(h->d_gs); cudaFree(h->d_hs);
  cudaFree(h->d_partial); cudaFree(h->d_out); cudaFree(h->d_summary); cudaFree(h->d_dbg);
  cudaFree(h->d_link); cudaFree(h->d_es_ptr); cudaFree(h->d_es_slot); cudaFree(h->d_es); cudaFree(h->d_En);
  if (h->h_x) cudaFreeHost(h->h_x);
  if (h->h_summary) cudaFreeHost(h->h_summary);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return OSB_OK;
}


// ---- host: path cover + node numbering for the chain preconditioner ------------------------------------------------
// Greedy maximum-weight path cover: factors sorted by information weight (descending, ties by index), a factor joins two
// free nodes when both still have degree < 2 and lie in different components (union-find: no cycles).  On a swarm graph
// this recovers every drone's odometry chain (ego-motion edges carry ~100x the information of loops / UWB).  Nodes are
// then numbered path by path (fixed nodes last); link[i] = 1 iff node i-1 precedes i on its path and i % 16 != 0.
struct ChainPlan {
  std::vector<int32_t> order;   // new id -> caller's id
  std::vector<int32_t> inv;     // caller's id -> new id
  std::vector<uint8_t> link;    // by new id
};

static double factor_weight(int type, const double* pl) {
  if (type == OSB_FACTOR_DISTANCE) return pl[1] * pl[1];
  if (type == OSB_FACTOR_RELPOSE) { double w = 0.0; for (int i = 4; i < 20; ++i) w += pl[i] * pl[i]; return w; }
  return pl[20] > 0.0 ? 1.0 / (pl[20] * pl[20]) : 0.0;
}

static void build_chain_plan(int n, const uint8_t* fixed, int m, const int32_t* type, const int32_t* ia, const int32_t* ib,
                             const double* payload, ChainPlan& pl) {
  // sort key: (inverted bits of the weight as a non-negative float) : factor index -> ascending u64 order is weight
  // descending, ties by index (weights only need float resolution here)
  std::vector<uint64_t> keys(m);
  for (int f = 0; f < m; ++f) {
    const float wf = (float)factor_weight(type[f], payload + (size_t)f * OSB_PAYLOAD_LEN);
    uint32_t bits;
    std::memcpy(&bits, &wf, sizeof(bits));
    keys[f] = ((uint64_t)(~bits) << 32) | (uint32_t)f;
  }
  std::sort(keys.begin(), keys.end());
  std::vector<int32_t> parent(n), nb0(n, -1), nb1(n, -1);
  for (int i = 0; i < n; ++i) parent[i] = i;
  auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
  for (uint64_t key : keys) {
    const int32_t f = (int32_t)(key & 0xffffffffu);
    const int a = ia[f], b = ib[f];
    if (fixed[a] || fixed[b] || nb1[a] >= 0 || nb1[b] >= 0) continue;
    const int ra = find(a), rb = find(b);
    if (ra == rb) continue;
    parent[ra] = rb;
    (nb0[a] < 0 ? nb0[a] : nb1[a]) = b;
    (nb0[b] < 0 ? nb0[b] : nb1[b]) = a;
  }
  pl.order.clear(); pl.order.reserve(n);
  pl.inv.assign(n, -1); pl.link.assign(n, 0);
  for (int s0 = 0; s0 < n; ++s0) {
    if (fixed[s0] || pl.inv[s0] >= 0 || nb1[s0] >= 0) continue;      // start at path ends (degree <= 1)
    int prev = -1, cur = s0;
    while (cur >= 0) {
      const int id = (int)pl.order.size();
      pl.inv[cur] = id; pl.order.push_back(cur);
      if (prev >= 0 && (id % 16) != 0) pl.link[id] = 1;
      const int nxt = (nb0[cur] >= 0 && nb0[cur] != prev) ? nb0[cur] : (nb1[cur] >= 0 && nb1[cur] != prev) ? nb1[cur] : -1;
      prev = cur; cur = nxt;
    }
  }
  for (int i = 0; i < n; ++i)
    if (pl.inv[i] < 0) { pl.inv[i] = (int)pl.order.size(); pl.order.push_back(i); }   // fixed nodes last
}

// host-only view of the plan (tests): order_out[new] = caller's node id, link_out[new]
extern "C" osb_status osb_solver_chain_plan(int n_nodes, const uint8_t* fixed, int n_factors, const int32_t* type,
                                            const int32_t* ia, const int32_t* ib, const double* payload,
                                            int32_t* order_out, uint8_t* link_out) {
  OSB_REQUIRE(fixed && type && ia && ib && payload && order_out && link_out && n_nodes > 0 && n_factors > 0, "bad argument");
  for (int f = 0; f < n_factors; ++f)
    OSB_REQUIRE(ia[f] >= 0 && ia[f] < n_nodes && ib[f] >= 0 && ib[f] < n_nodes && ia[f] != ib[f], "bad factor indices");
  ChainPlan pl;
  build_chain_plan(n_nodes, fixed, n_factors, type, ia, ib, payload, pl);
  for (int i = 0; i < n_nodes; ++i) { order_out[i] = pl.order[i]; link_out[i] = pl.link[i]; }
  return OSB_OK;
}

static osb_status validate_graph(int n_nodes, int n_factors, const int32_t* type, const int32_t* ia, const int32_t* ib) {
  for (int f = 0; f < n_factors; ++f) {
    OSB_REQUIRE(type[f] >= 0 && type[f] <= 2, "unknown factor type");
    OSB_REQUIRE(ia[f] >= 0 && ia[f] < n_nodes && ib[f] >= 0 && ib[f] < n_nodes, "factor node index out of range");
    // the reference skips factors whose two parameter blocks coincide (solver.cpp:1071-1073,1176)
    OSB_REQUIRE(ia[f] != ib[f], "factor connects a pose block to itself (the adapter must skip it)");
  }
  return OSB_OK;
}

// the solve proper.  upload_static: copy type / huber / payload to the device (one-shot entry point); the resident entry
// point has already appended them.  The caller holds h->mu.
static osb_status solver_run(osb_solver* h, int n_nodes, double* poses, const uint8_t* fixed, int n_factors,
                             const int32_t* type, const int32_t* ia, const int32_t* ib, const double* payload,
                             const uint8_t* huber, bool upload_static, const osb_solve_options* opt,
                             osb_solve_summary* summary, unsigned long long topo_key = 0) {
  osb_solve_options o;
  if (opt) o = *opt; else osb_solve_default_options(&o);
  // Internal node numbering: the paths of the chain plan are runs of consecutive ids (fixed nodes last).  Everything on
  // the device uses the internal ids; poses are permuted on the way in and out.
  const size_t n = n_nodes, m = n_factors;
  cudaStream_t st = h->stream;
  std::vector<double> x_p(4 * n);
  // resident graph with unchanged topology: plan, numbering and every index table are already on the device
  const bool reuse = topo_key != 0 && topo_key == h->cached_topo && h->cached_order.size() == n;
  std::vector<int32_t> order_local;
  int n_res = 0;
  if (reuse) {
    for (size_t i = 0; i < n; ++i) {
      const int o2 = h->cached_order[i];
      for (int k = 0; k < 4; ++k) x_p[4 * i + k] = poses[4 * (size_t)o2 + k];
    }
    n_res = h->cached_nres;
    memcpy(h->h_x, x_p.data(), 4 * n * sizeof(double));
    OSB_CUDA(cudaMemcpyAsync(h->d_x0, h->h_x, 4 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
  ChainPlan plan;
  build_chain_plan(n_nodes, fixed, n_factors, type, ia, ib, payload, plan);
  std::vector<int32_t> ia_p(m), ib_p(m);
  std::vector<uint8_t> fixed_p(n);
  for (size_t f = 0; f < m; ++f) { ia_p[f] = plan.inv[ia[f]]; ib_p[f] = plan.inv[ib[f]]; }
  for (size_t i = 0; i < n; ++i) {
    const int o = plan.order[i];
    fixed_p[i] = fixed[o];
    for (int k = 0; k < 4; ++k) x_p[4 * i + k] = poses[4 * (size_t)o + k];
  }
  // chain couplings: factor f couples node hi with hi-1 when its two nodes are consecutive and linked
  std::vector<int32_t> es_ptr(n + 1, 0), es_slot(m, -1);
  for (size_t f = 0; f < m; ++f) {
    const int a = ia_p[f], b = ib_p[f], hi = std::max(a, b);
    if (std::abs(a - b) == 1 && plan.link[hi]) es_ptr[hi + 1]++;
  }
  for (size_t i = 0; i < n; ++i) es_ptr[i + 1] += es_ptr[i];
  {
    std::vector<int32_t> fill(es_ptr.begin(), es_ptr.end() - 1);
    for (size_t f = 0; f < m; ++f) {
      const int a = ia_p[f], b = ib_p[f], hi = std::max(a, b);
      if (std::abs(a - b) == 1 && plan.link[hi]) es_slot[f] = 2 * (fill[hi]++) + (a == hi ? 1 : 0);
    }
  }
  // CSR of contribution slots: node n owns slots [ptr[n], ptr[n+1]); factors in index order within a node, so the
  // gather order -- and therefore every floating-point sum -- is fixed.
  std::vector<int32_t> ptr(n + 1, 0), slot_a(m), slot_b(m);
  for (size_t f = 0; f < m; ++f) { ptr[ia_p[f] + 1]++; ptr[ib_p[f] + 1]++; }
  for (size_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  {
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (size_t f = 0; f < m; ++f) { slot_a[f] = fill[ia_p[f]]++; slot_b[f] = fill[ib_p[f]]++; }
  }
  for (size_t f = 0; f < m; ++f)
    n_res += type[f] == OSB_FACTOR_DISTANCE ? 1 : type[f] == OSB_FACTOR_RELPOSE ? 4
             : (((int)payload[f * OSB_PAYLOAD_LEN + 10] & 1) ? 3 : 2);
  OSB_CUDA(cudaMemcpyAsync(h->d_fixed, fixed_p.data(), n, cudaMemcpyHostToDevice, st));
  if (upload_static) {
    OSB_CUDA(cudaMemcpyAsync(h->d_huber, huber, m, cudaMemcpyHostToDevice, st));
    OSB_CUDA(cudaMemcpyAsync(h->d_type, type, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    OSB_CUDA(cudaMemcpyAsync(h->d_payload, payload, m * OSB_PAYLOAD_LEN * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  OSB_CUDA(cudaMemcpyAsync(h->d_ia, ia_p.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_ib, ib_p.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_ptr, ptr.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_slot_a, slot_a.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_slot_b, slot_b.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_x0, x_p.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_link, plan.link.data(), n, cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_es_ptr, es_ptr.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_es_slot, es_slot.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  // the staging vectors above die at the end of this block: the copies must have left them
  OSB_CUDA(cudaStreamSynchronize(st));
  h->cached_order = plan.order;
  h->cached_nres = n_res;
  h->cached_topo = topo_key;               // 0 (one-shot solve) never matches
  }
  const std::vector<int32_t>& order = h->cached_order;

  SolverDev P;
  P.n = n_nodes; P.m = n_factors;
  P.fixed = h->d_fixed; P.ftype = h->d_type; P.ia = h->d_ia; P.ib = h->d_ib; P.huber = h->d_huber;
  P.payload = h->d_payload; P.node_ptr = h->d_ptr; P.slot_a = h->d_slot_a; P.slot_b = h->d_slot_b;
  P.x[0] = h->d_x0; P.x[1] = h->d_x1; P.Jg = h->d_Jg;
  double* nv = h->d_nodevec;
  const size_t N4 = 4 * (size_t)h->max_nodes, N16 = 16 * (size_t)h->max_nodes;
  P.g = nv; P.D = nv + N4; P.p = nv + 2 * N4; P.z = nv + 3 * N4; P.res = nv + 4 * N4; P.Ap = nv + 5 * N4;
  P.delta = nv + 6 * N4; P.Hnn = nv + 7 * N4; P.Minv = nv + 7 * N4 + N16;
  P.cs = h->d_cs; P.gs = h->d_gs; P.hs = h->d_hs; P.partial = h->d_partial; P.opt = o; P.summary = h->d_summary; P.poses_out = h->d_out; P.dbg = h->d_dbg;
  P.use_chain = 0; P.link = h->d_link; P.es_ptr = h->d_es_ptr; P.es_slot = h->d_es_slot; P.es = h->d_es; P.En = h->d_En;

  // inner precision: fp32 PCG unless the caller asks for a tighter inner solve than fp32 can deliver
  const bool f32 = o.inner_precision == OSB_INNER_FP32 || (o.inner_precision == OSB_INNER_AUTO && o.pcg_tolerance >= 1e-4);
  const size_t tsz = f32 ? sizeof(float) : sizeof(double);
  const void* kern = f32 ? (const void*)graph_solve_kernel<float> : (const void*)graph_solve_kernel<double>;
  // launch shape: ONE thread-block cluster (hardware barrier, ~0.2 us) when the factor list fits 16 CTAs with their
  // Jacobians in shared memory; otherwise a cooperative grid (software grid barrier).
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(GS_THREADS); cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
  bool launched_cluster = false;
  {
    const int G = std::max(1, std::min(GS_MAX_CLUSTER, cdiv(std::max(n_nodes, n_factors), GS_THREADS)));
    const int fpc = cdiv(n_factors, G);
    const size_t jbytes = (size_t)fpc * 32 * tsz, cbytes = (size_t)16 * GS_THREADS * tsz;
    if (h->cluster_ok && jbytes <= (size_t)GS_SMEM_J_MAX && cdiv(n_nodes, GS_THREADS) <= 4 * G) {
      // chain preconditioner: needs the fast path (one thread per node, <= GS_KF factors per thread) and room for L
      const bool chain = o.preconditioner != OSB_PRECOND_BLOCK_JACOBI && fpc <= GS_KF * GS_THREADS &&
                         n_nodes <= G * GS_THREADS && jbytes + cbytes <= (size_t)GS_SMEM_DYN_MAX;
      P.use_chain = chain ? 1 : 0;
      cfg.gridDim = dim3(G); cfg.dynamicSmemBytes = jbytes + (chain ? cbytes : 0);
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) {
        P.fpc = fpc; P.use_cluster = 1; P.j_in_smem = 1;
        launched_cluster = true;
      } else {
        cudaGetLastError();
        P.use_chain = 0;
      }
    }
  }
  if (!launched_cluster) {
    P.use_chain = 0;
    int per_sm = 0;
    const int G0 = std::max(1, std::min(num_sms(), cdiv(std::max(n_nodes, n_factors), GS_THREADS)));
    int fpc = cdiv(n_factors, G0);
    const size_t jbytes = (size_t)fpc * 32 * tsz;
    P.j_in_smem = (jbytes <= (size_t)GS_SMEM_J_MAX) ? 1 : 0;
    const size_t smem = P.j_in_smem ? jbytes : 0;
    OSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GS_THREADS, smem));
    if (per_sm < 1) { set_error("osb_solver_solve", "solve kernel cannot be made resident"); return OSB_ERR_CUDA; }
    P.fpc = fpc; P.use_cluster = 0;
    cfg.gridDim = dim3(G0); cfg.dynamicSmemBytes = smem;
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
  }
  h->last_grid = (int)cfg.gridDim.x; h->last_cluster = P.use_cluster; h->last_jsmem = P.j_in_smem; h->last_chain = P.use_chain;
  h->last_f32 = f32 ? 1 : 0;
  OSB_CUDA(cudaEventRecord(h->ev0, st));
  {
    void* kargs[1] = {&P};
    OSB_CUDA(cudaLaunchKernelExC(&cfg, kern, kargs));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  OSB_CUDA(cudaEventRecord(h->ev1, st));
  OSB_CUDA(cudaMemcpyAsync(h->h_x, h->d_out, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(h->h_summary, h->d_summary, sizeof(osb_solve_summary), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  memcpy(x_p.data(), h->h_x, 4 * n * sizeof(double));
  *summary = *h->h_summary;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) poses[4 * (size_t)order[i] + k] = x_p[4 * i + k];
  float ms = 0.f;
  OSB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  summary->solve_ms = ms;
  summary->n_residuals = n_res;
  return OSB_OK;
}

extern "C" osb_status osb_solver_solve(osb_solver* h, int n_nodes, double* poses, const uint8_t* fixed, int n_factors,
                                       const int32_t* type, const int32_t* ia, const int32_t* ib,
                                       const double* payload, const uint8_t* huber, const osb_solve_options* opt,
                                       osb_solve_summary* summary) {
  OSB_REQUIRE(h && poses && fixed && type && ia && ib && payload && huber && summary, "null argument");
  OSB_REQUIRE(n_nodes > 0 && n_nodes <= h->max_nodes && n_factors > 0 && n_factors <= h->max_factors,
              "graph larger than the solver capacity");
  osb_status s = validate_graph(n_nodes, n_factors, type, ia, ib);
  if (s != OSB_OK) return s;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  h->g_static_valid = false;              // the device factor arrays now hold this graph, not the resident one
  h->cached_topo = 0;                     // ... and so do the index tables
  return solver_run(h, n_nodes, poses, fixed, n_factors, type, ia, ib, payload, huber, true, opt, summary);
}

// -------------------------------------------------------------------------------------------------------------
// Resident graph.  The reference re-flattens its whole window into a ceres::Problem on every
// solve (setup_problem_with_sferror / _loops_and_detections / _ego_motion, swarm_localization_solver.cpp:1064-1214).
// Here the factor list lives in device memory between solves: add_new_swarm_frame / add_new_loop_connection
// (swarm_localization_solver.hpp:197-214) append what a frame adds, only the new factors cross PCIe, and the poses
// stay where the last solve left them (the reference's est_poses persist the same way).
// -------------------------------------------------------------------------------------------------------------
extern "C" osb_status osb_solver_graph_clear(osb_solver* h) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  h->g_poses.clear(); h->g_payload.clear(); h->g_fixed.clear(); h->g_huber.clear();
  h->g_type.clear(); h->g_ia.clear(); h->g_ib.clear();
  h->g_uploaded = 0; h->g_static_valid = true;
  ++h->g_topo_version;
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_add_nodes(osb_solver* h, int n, const double* poses, const uint8_t* fixed,
                                                 int32_t* first_id) {
  OSB_REQUIRE(h != nullptr && n > 0 && poses != nullptr, "bad argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  const size_t have = h->g_fixed.size();
  if (have + (size_t)n > (size_t)h->max_nodes) { set_error("osb_solver_graph_add_nodes", "node capacity exceeded"); return OSB_ERR_CAPACITY; }
  if (first_id) *first_id = (int32_t)have;
  h->g_poses.insert(h->g_poses.end(), poses, poses + 4 * (size_t)n);
  for (int i = 0; i < n; ++i) h->g_fixed.push_back(fixed ? fixed[i] : 0);
  ++h->g_topo_version;
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_add_factors(osb_solver* h, int m, const int32_t* type, const int32_t* ia,
                                                   const int32_t* ib, const double* payload, const uint8_t* huber) {
  OSB_REQUIRE(h != nullptr && m > 0 && type && ia && ib && payload && huber, "bad argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  const size_t have = h->g_type.size();
  if (have + (size_t)m > (size_t)h->max_factors) { set_error("osb_solver_graph_add_factors", "factor capacity exceeded"); return OSB_ERR_CAPACITY; }
  osb_status s = validate_graph((int)h->g_fixed.size(), m, type, ia, ib);
  if (s != OSB_OK) return s;
  h->g_type.insert(h->g_type.end(), type, type + m);
  h->g_ia.insert(h->g_ia.end(), ia, ia + m);
  h->g_ib.insert(h->g_ib.end(), ib, ib + m);
  h->g_huber.insert(h->g_huber.end(), huber, huber + m);
  h->g_payload.insert(h->g_payload.end(), payload, payload + (size_t)m * OSB_PAYLOAD_LEN);
  ++h->g_topo_version;
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_set_fixed(osb_solver* h, int node, int fixed) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  OSB_REQUIRE(node >= 0 && (size_t)node < h->g_fixed.size(), "node out of range");
  h->g_fixed[node] = fixed ? 1 : 0;
  ++h->g_topo_version;
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_set_poses(osb_solver* h, int first, int n, const double* poses) {
  OSB_REQUIRE(h != nullptr && poses != nullptr, "bad argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  OSB_REQUIRE(first >= 0 && n >= 0 && (size_t)(first + n) <= h->g_fixed.size(), "node range out of bounds");
  std::copy(poses, poses + 4 * (size_t)n, h->g_poses.begin() + 4 * (size_t)first);
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_get_poses(osb_solver* h, int first, int n, double* poses) {
  OSB_REQUIRE(h != nullptr && poses != nullptr, "bad argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  OSB_REQUIRE(first >= 0 && n >= 0 && (size_t)(first + n) <= h->g_fixed.size(), "node range out of bounds");
  std::copy(h->g_poses.begin() + 4 * (size_t)first, h->g_poses.begin() + 4 * (size_t)(first + n), poses);
  return OSB_OK;
}

extern "C" osb_status osb_solver_graph_size(osb_solver* h, int32_t* n_nodes, int32_t* n_factors) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  if (n_nodes) *n_nodes = (int32_t)h->g_fixed.size();
  if (n_factors) *n_factors = (int32_t)h->g_type.size();
  return OSB_OK;
}

// sliding window (solver.cpp:186-202 trims old keyframes): drop the first `n_nodes` nodes and every factor touching them;
// the remaining nodes are renumbered (id - n_nodes).  A rare operation: the device factor arrays are rebuilt.
extern "C" osb_status osb_solver_graph_drop_oldest(osb_solver* h, int n_nodes) {
  OSB_REQUIRE(h != nullptr && n_nodes >= 0, "bad argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  OSB_REQUIRE((size_t)n_nodes <= h->g_fixed.size(), "cannot drop more nodes than the graph holds");
  if (n_nodes == 0) return OSB_OK;
  h->g_poses.erase(h->g_poses.begin(), h->g_poses.begin() + 4 * (size_t)n_nodes);
  h->g_fixed.erase(h->g_fixed.begin(), h->g_fixed.begin() + n_nodes);
  size_t w = 0;
  for (size_t f = 0; f < h->g_type.size(); ++f) {
    if (h->g_ia[f] < n_nodes || h->g_ib[f] < n_nodes) continue;
    h->g_type[w] = h->g_type[f]; h->g_ia[w] = h->g_ia[f] - n_nodes; h->g_ib[w] = h->g_ib[f] - n_nodes;
    h->g_huber[w] = h->g_huber[f];
    if (w != f) std::copy(h->g_payload.begin() + f * OSB_PAYLOAD_LEN, h->g_payload.begin() + (f + 1) * OSB_PAYLOAD_LEN,
                          h->g_payload.begin() + w * OSB_PAYLOAD_LEN);
    ++w;
  }
  h->g_type.resize(w); h->g_ia.resize(w); h->g_ib.resize(w); h->g_huber.resize(w); h->g_payload.resize(w * OSB_PAYLOAD_LEN);
  h->g_uploaded = 0;                       // factor positions moved: re-send on the next solve
  ++h->g_topo_version;
  return OSB_OK;
}

extern "C" osb_status osb_solver_solve_resident(osb_solver* h, const osb_solve_options* opt, osb_solve_summary* summary) {
  OSB_REQUIRE(h != nullptr && summary != nullptr, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  const size_t n = h->g_fixed.size(), m = h->g_type.size();
  OSB_REQUIRE(n > 0 && m > 0, "the resident graph is empty");
  if (!h->g_static_valid) { h->g_uploaded = 0; h->g_static_valid = true; h->cached_topo = 0; }
  if (h->g_uploaded < m) {                 // only the factors added since the last solve cross PCIe
    const size_t f0 = h->g_uploaded, k = m - f0;
    cudaStream_t st = h->stream;
    OSB_CUDA(cudaMemcpyAsync(h->d_huber + f0, h->g_huber.data() + f0, k, cudaMemcpyHostToDevice, st));
    OSB_CUDA(cudaMemcpyAsync(h->d_type + f0, h->g_type.data() + f0, k * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    OSB_CUDA(cudaMemcpyAsync(h->d_payload + f0 * OSB_PAYLOAD_LEN, h->g_payload.data() + f0 * OSB_PAYLOAD_LEN,
                             k * OSB_PAYLOAD_LEN * sizeof(double), cudaMemcpyHostToDevice, st));
    h->g_uploaded = m;
  }
  return solver_run(h, (int)n, h->g_poses.data(), h->g_fixed.data(), (int)m, h->g_type.data(), h->g_ia.data(),
                    h->g_ib.data(), h->g_payload.data(), h->g_huber.data(), false, opt, summary, h->g_topo_version);
}

extern "C" osb_status osb_solver_phase_cycles(osb_solver* h, double* out12) {
  OSB_REQUIRE(h != nullptr && out12 != nullptr, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  long long c[8];
  OSB_CUDA(cudaMemcpy(c, h->d_dbg, sizeof(c), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) out12[i] = (double)c[i];
  out12[8] = h->last_grid; out12[9] = h->last_cluster; out12[10] = h->last_jsmem + 2 * h->last_chain + 4 * h->last_f32; out12[11] = GS_THREADS;
  return OSB_OK;
}

// profiling aid: per-warp cycles spent in the chain preconditioner sweeps of the last solve, [16 CTAs][8 warps]
extern "C" osb_status osb_solver_chain_cycles(osb_solver* h, double* out128) {
  OSB_REQUIRE(h != nullptr && out128 != nullptr, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  long long c[128];
  OSB_CUDA(cudaMemcpy(c, h->d_dbg + 8, sizeof(c), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 128; ++i) out128[i] = (double)c[i];
  return OSB_OK;
}

extern "C" osb_status osb_solver_linearize(osb_solver* h, int n_nodes, const double* poses, int n_factors,
                                           const int32_t* type, const int32_t* ia, const int32_t* ib,
                                           const double* payload, double* r, double* Ja, double* Jb) {
  OSB_REQUIRE(h && poses && type && ia && ib && payload && r && Ja && Jb, "null argument");
  OSB_REQUIRE(n_nodes > 0 && n_nodes <= h->max_nodes && n_factors > 0 && n_factors <= h->max_factors,
              "graph larger than the solver capacity");
  osb_status s = validate_graph(n_nodes, n_factors, type, ia, ib);
  if (s != OSB_OK) return s;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  const size_t n = n_nodes, m = n_factors;
  cudaStream_t st = h->stream;
  OSB_CUDA(cudaMemcpyAsync(h->d_type, type, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_ia, ia, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_ib, ib, m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_payload, payload, m * OSB_PAYLOAD_LEN * sizeof(double), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_x0, poses, 4 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  double* d_r = h->d_lin;             // [m][4]
  double* d_ja = h->d_lin + 4 * m;    // [m][16]
  double* d_jb = h->d_lin + 20 * m;   // [m][16]
  OSB_LAUNCH(graph_linearize_kernel, cdiv(n_factors, 128), 128, 0, st, n_factors, h->d_x0, h->d_type, h->d_ia, h->d_ib,
             h->d_payload, d_r, d_ja, d_jb);
  OSB_CHECK_LAUNCH();
  OSB_CUDA(cudaMemcpyAsync(r, d_r, 4 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(Ja, d_ja, 16 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(Jb, d_jb, 16 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  return OSB_OK;
}
